"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/spk_oracle.c (the CPU restatement of the
reference's numeric path).  Import only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libspkoracle.so")
_lib = None
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "spk_oracle.c")
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(src) > os.path.getmtime(_LIBPATH):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libspkoracle.so"])
    return _LIBPATH


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        i = C.c_int64
        L.spko_lufactor.argtypes = [i, i, I64P, I64P, I64P, I64P, I64P, F64P, I64P, F64P, I64P]
        L.spko_lufactor.restype = i
        L.spko_lulsolve.argtypes = [i, I64P, I64P, I64P, I64P, F64P, I64P, F64P]
        L.spko_luusolve.argtypes = [i, i, I64P, I64P, I64P, I64P, F64P, I64P, F64P, F64P]
        L.spko_ldltfactor.argtypes = [i, i, I64P, I64P, I64P, I64P, I64P, F64P]
        L.spko_ldltfactor.restype = i
        L.spko_ldltsolve.argtypes = [i, I64P, I64P, I64P, I64P, F64P, F64P]
        L.spko_schedule_stats.argtypes = [i, i, I64P, I64P, I64P, I64P, F64P]
        L.spko_use_blas.argtypes = [C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


def use_openblas(enable=True):
    """Route the oracle's dense call sites to the OpenBLAS bundled with SciPy (LP64, `scipy_` prefix) —
    the library family Julia's libblastrampoline forwards to (SpkSpdMMOps.jl:17-21)."""
    if not enable:
        return lib().spko_use_blas(None, None)
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
    if not cands:
        raise RuntimeError("SciPy's OpenBLAS not found")
    rc = lib().spko_use_blas(os.path.abspath(cands[0]).encode(), b"scipy_")
    if rc != 0:
        raise RuntimeError(f"spko_use_blas failed: {rc}")
    return 0


def lufactor(s, lnz, unz, ipiv):
    """Oracle of `_lufactor!` on the arrays of a `_SparseBase`-like object; in place."""
    return int(lib().spko_lufactor(s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, s.xlnz, lnz, s.xunz, unz, ipiv))


def lusolve(s, lnz, unz, ipiv, rhs):
    """`_lulsolve!` then `_luusolve!` on a permuted rhs, in place."""
    L = lib()
    L.spko_lulsolve(s.nsuper, s.xsuper, s.xlindx, s.lindx, s.xlnz, lnz, ipiv, rhs)
    L.spko_luusolve(s.n, s.nsuper, s.xsuper, s.xlindx, s.lindx, s.xlnz, lnz, s.xunz, unz, rhs)
    return rhs


def ldltfactor(s, lnz):
    return int(lib().spko_ldltfactor(s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, s.xlnz, lnz))


def ldltsolve(s, lnz, rhs):
    lib().spko_ldltsolve(s.nsuper, s.xsuper, s.xlindx, s.lindx, s.xlnz, lnz, rhs)
    return rhs


def triangularsolve(s, lnz, unz, ipiv, b):
    """`_triangularsolve!` (SpkSparseBase.jl:400-416 / SpkSparseSpdBase.jl:334-356) with the oracle."""
    rhs = np.ascontiguousarray(b[s.order.rperm - 1], dtype=np.float64)
    if s.spd:
        ldltsolve(s, lnz, rhs)
    else:
        lusolve(s, lnz, unz, ipiv, rhs)
    return rhs[s.order.rinvp - 1]


def schedule_stats(s):
    out = np.zeros(4)
    lib().spko_schedule_stats(s.n, s.nsuper, s.xsuper, s.snode, s.xlindx, s.lindx, out)
    return dict(ncmod=out[0], nrank1=out[1], flops_lu=out[2], flops_spd=out[3])
