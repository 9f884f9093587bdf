"""The C-ABI shared library loads and exports every symbol include/spk_b200.h declares
(no compute calls: there is no GPU in the CPU test run)."""
import ctypes
import os
import re

import pytest

import sparspak_jl_b200 as spk
from sparspak_jl_b200 import build, _cudalib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spk_b200.h")).read()
    return sorted(set(re.findall(r"SPK_API\s+[\w\s\*]+?\b(spk_\w+)\s*\(", hdr)))


def test_library_builds_and_exports_all_declared_symbols():
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in spk_b200.h but not exported"
    assert sorted(_cudalib.SYMBOLS) == names


def test_version_and_device_count_callable_without_gpu():
    L = _cudalib.lib()
    assert b"sm_100a" in L.spk_version()
    assert L.spk_device_count() >= 0


def test_no_cpu_fallback_without_gpu():
    """Without a GPU the product path must fail loudly, not fall back to the oracle / CPU."""
    L = _cudalib.lib()
    if L.spk_device_count() > 0:
        pytest.skip("GPU present")
    s = spk.SparseSolver(spk.matrices.laplacian2d(4))
    spk.findorder(s); spk.symbolicfactor(s); spk.inmatrix(s)
    with pytest.raises(_cudalib.SpkError):
        spk.factor(s)


def test_product_sources_do_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "sparspak.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "spko_" not in txt and "import oracle" not in txt and "libspkoracle" not in txt, f


def test_julia_shim_binds_only_declared_symbols_with_matching_arity():
    """julia/SparspakB200.jl cannot run here (no Julia in the image): at least every `ccall((:sym, libspk), ...)` in
    it must name a symbol the header declares, with as many argument types as the C prototype has parameters."""
    hdr = open(os.path.join(ROOT, "include", "spk_b200.h")).read()
    proto = {}
    for m in re.finditer(r"SPK_API\s+[\w\s\*]+?\b(spk_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        args = m.group(2).strip()
        proto[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    jl = open(os.path.join(ROOT, "julia", "SparspakB200.jl")).read()
    calls = re.findall(r"ccall\(\(:(spk_\w+), libspk\),\s*[\w{}\.]+,\s*\(([^()]*)\)", jl, re.S)
    assert len(calls) >= 20
    for name, argt in calls:
        assert name in proto, f"{name} bound in the Julia shim but not declared in spk_b200.h"
        n = len([a for a in argt.replace("\n", " ").split(",") if a.strip()])
        assert n == proto[name], f"{name}: {n} ccall argument types, {proto[name]} C parameters"
