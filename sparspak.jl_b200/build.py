"""In-tree builds: the CUDA C-ABI library (sm_100a) and the host structure library.  Built artefacts are git-ignored but travel to the GPU box with the
gpurun snapshot, so nothing is JIT-compiled there."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
HOST = os.path.join(PKG_DIR, "host")
INCLUDE = os.path.join(ROOT, "include")

LIB_CUDA = os.path.join(CSRC, "libspkb200.so")
LIB_HOST = os.path.join(HOST, "libspkhost.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1]))
    return r.stdout + r.stderr


def build_host(force=False):
    src = [os.path.join(HOST, "spk_host.cpp")]
    if force or _stale(LIB_HOST, src):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-o", LIB_HOST] + src)
    return LIB_HOST


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_cuda(force=False, verbose=False):
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))]
    deps += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    if not (force or _stale(LIB_CUDA, deps)):
        return LIB_CUDA
    if shutil.which(NVCC) is None and not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found; cannot build libspkb200.so")
    cmd = [NVCC] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-shared", "-o", LIB_CUDA] + srcs + ["-lcudart", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    out = _run(cmd)
    if verbose:
        print(out)
    return LIB_CUDA


def build_all(force=False):
    build_host(force)
    build_cuda(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB_HOST, LIB_CUDA)
