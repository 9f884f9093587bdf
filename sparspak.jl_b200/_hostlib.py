"""ctypes bindings of the host structure library (host/spk_host.cpp)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB_HOST
        if not os.path.exists(path):
            path = _build.build_host()
        L = C.CDLL(path)
        i = C.c_int64
        L.spkh_mmd.argtypes = [i, I64P, I64P, I64P, I64P]
        L.spkh_nd_grid.argtypes = [i, i, i, i, i, I64P, I64P]
        L.spkh_etree.argtypes = [i, I64P, I64P, I64P, I64P, I64P]
        L.spkh_postorder.argtypes = [i, I64P, I64P, I64P, C.c_void_p]
        L.spkh_colcounts.argtypes = [i, I64P, I64P, I64P, I64P, I64P, I64P]
        L.spkh_colcounts.restype = i
        L.spkh_findsupernodes.argtypes = [i, I64P, I64P, i, I64P, I64P, I64P]
        L.spkh_nonzeroindexs.argtypes = [i, I64P, i, I64P, I64P, C.c_void_p]
        L.spkh_symbolicfact.argtypes = [i, I64P, I64P, I64P, I64P, I64P, i, I64P, I64P, I64P, I64P]
        L.spkh_inmatrix_map_lu.argtypes = [i] + [I64P] * 11
        L.spkh_inmatrix_map_lu.restype = i
        L.spkh_inmatrix_map_spd.argtypes = [i] + [I64P] * 10
        L.spkh_inmatrix_map_spd.restype = i
        L.spkh_scatter_values.argtypes = [i, I64P, F64P, F64P, C.c_void_p]
        L.spkh_workcounts.argtypes = [i, I64P, F64P]
        _lib = L
    return _lib
