import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import sparspak_jl_b200 as spk
        from sparspak_jl_b200 import _cudalib
        return _cudalib.lib().spk_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_cpu_libs():
    """Host structure library, oracle and the host simulator of the device schedule."""
    import sparspak_jl_b200 as spk
    from sparspak_jl_b200 import build
    build.build_host()
    import oracle
    oracle.build()
    hs = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(hs, "libhostsim.so")
    srcs = [os.path.join(hs, "hostsim.cpp"), os.path.join(build.CSRC, "plan.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
                               "-I", build.CSRC, "-o", so, srcs[0]])
    yield
