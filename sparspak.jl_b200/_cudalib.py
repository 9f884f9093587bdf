"""ctypes binding of the CUDA C-ABI library (csrc/libspkb200.so, include/spk_b200.h).

The numeric path has NO fallback: if the library is missing or there is no GPU, creating a
`Plan` raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

SYMBOLS = [
    "spk_lufactor_f64", "spk_lulsolve_f64", "spk_luusolve_f64", "spk_ldltfactor_f64", "spk_ldltsolve_f64",
    "spk_lufactor_f32", "spk_lulsolve_f32", "spk_luusolve_f32", "spk_ldltfactor_f32", "spk_ldltsolve_f32",
    "spk_plan_inmatrix_f32", "spk_plan_get_factors_f32", "spk_plan_triangularsolve_f32",
    "spk_plan_set_matrix", "spk_plan_residual", "spk_plan_refine",
    "spk_plan_create", "spk_plan_destroy", "spk_plan_inmatrix", "spk_plan_reassemble", "spk_plan_set_values", "spk_plan_factor",
    "spk_plan_get_factors", "spk_plan_set_factors", "spk_plan_solve", "spk_plan_set_perm",
    "spk_plan_triangularsolve", "spk_plan_device_ptr", "spk_plan_device_len", "spk_plan_factor_phase",
    "spk_plan_solve_device", "spk_plan_solve_phase", "spk_nccl_unique_id", "spk_plan_comm_init", "spk_plan_factor_multi", "spk_plan_solve_multi",
    "spk_multi_create", "spk_multi_destroy", "spk_multi_plan", "spk_multi_inmatrix", "spk_multi_set_values", "spk_multi_factor",
    "spk_multi_get_factors", "spk_multi_set_perm", "spk_multi_triangularsolve", "spk_plan_xchg_info", "spk_plan_stat", "spk_plan_statf", "spk_plan_condest", "spk_cache_clear", "spk_last_error", "spk_device_count", "spk_version",
]


class SpkError(RuntimeError):
    pass


def lib():
    """Load libspkb200.so (built in-tree by build.py).  Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_CUDA
    if not os.path.exists(path):
        raise SpkError(f"{path} not found: run `python __graft_entry__.py` (build()) first; "
                       "the numeric path has no CPU fallback")
    L = C.CDLL(path)
    i64, i32, vp, dbl = C.c_int64, C.c_int32, C.c_void_p, C.c_double
    L.spk_lufactor_f64.argtypes = [i64, i64, I64P, I64P, I64P, I64P, I64P, F64P, I64P, F64P, I64P]
    L.spk_lufactor_f64.restype = i64
    L.spk_lulsolve_f64.argtypes = [i64, I64P, I64P, I64P, I64P, F64P, I64P, F64P]
    L.spk_lulsolve_f64.restype = i64
    L.spk_luusolve_f64.argtypes = [i64, i64, I64P, I64P, I64P, I64P, F64P, I64P, F64P, F64P]
    L.spk_luusolve_f64.restype = i64
    L.spk_ldltfactor_f64.argtypes = [i64, i64, I64P, I64P, I64P, I64P, I64P, F64P]
    L.spk_ldltfactor_f64.restype = i64
    L.spk_ldltsolve_f64.argtypes = [i64, I64P, I64P, I64P, I64P, F64P, F64P]
    L.spk_ldltsolve_f64.restype = i64
    F32P = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    L.spk_lufactor_f32.argtypes = [i64, i64, I64P, I64P, I64P, I64P, I64P, F32P, I64P, F32P, I64P]
    L.spk_lufactor_f32.restype = i64
    L.spk_lulsolve_f32.argtypes = [i64, I64P, I64P, I64P, I64P, F32P, I64P, F32P]
    L.spk_lulsolve_f32.restype = i64
    L.spk_luusolve_f32.argtypes = [i64, i64, I64P, I64P, I64P, I64P, F32P, I64P, F32P, F32P]
    L.spk_luusolve_f32.restype = i64
    L.spk_ldltfactor_f32.argtypes = [i64, i64, I64P, I64P, I64P, I64P, I64P, F32P]
    L.spk_ldltfactor_f32.restype = i64
    L.spk_ldltsolve_f32.argtypes = [i64, I64P, I64P, I64P, I64P, F32P, F32P]
    L.spk_ldltsolve_f32.restype = i64
    L.spk_plan_inmatrix_f32.argtypes = [vp, i64, vp, F32P]
    L.spk_plan_inmatrix_f32.restype = i64
    L.spk_plan_get_factors_f32.argtypes = [vp, vp, vp, vp]
    L.spk_plan_get_factors_f32.restype = i64
    L.spk_plan_triangularsolve_f32.argtypes = [vp, F32P, i64, i64]
    L.spk_plan_triangularsolve_f32.restype = i64
    L.spk_plan_set_matrix.argtypes = [vp, i64, I64P, I64P, F64P]
    L.spk_plan_set_matrix.restype = i64
    L.spk_plan_residual.argtypes = [vp, vp, vp, i64, i64, vp, vp]
    L.spk_plan_residual.restype = i64
    L.spk_plan_refine.argtypes = [vp, vp, vp, i64, i64, i32, dbl, vp]
    L.spk_plan_refine.restype = i64
    L.spk_plan_create.argtypes = [i64, i64, I64P, I64P, I64P, I64P, I64P, vp, i32, i32, i32]
    L.spk_plan_create.restype = vp
    L.spk_plan_destroy.argtypes = [vp]
    L.spk_plan_destroy.restype = None
    L.spk_plan_inmatrix.argtypes = [vp, i64, vp, F64P]
    L.spk_plan_inmatrix.restype = i64
    L.spk_plan_reassemble.argtypes = [vp]
    L.spk_plan_reassemble.restype = i64
    L.spk_plan_set_values.argtypes = [vp, F64P, vp]
    L.spk_plan_set_values.restype = i64
    L.spk_plan_factor.argtypes = [vp]
    L.spk_plan_factor.restype = i64
    L.spk_plan_get_factors.argtypes = [vp, vp, vp, vp]
    L.spk_plan_get_factors.restype = i64
    L.spk_plan_set_factors.argtypes = [vp, F64P, vp, vp]
    L.spk_plan_set_factors.restype = i64
    L.spk_plan_solve.argtypes = [vp, F64P, i64, i64, i32]
    L.spk_plan_solve.restype = i64
    L.spk_plan_set_perm.argtypes = [vp, I64P, I64P]
    L.spk_plan_set_perm.restype = i64
    L.spk_plan_triangularsolve.argtypes = [vp, F64P, i64, i64]
    L.spk_plan_triangularsolve.restype = i64
    L.spk_plan_device_ptr.argtypes = [vp, i32]
    L.spk_plan_device_ptr.restype = vp
    L.spk_plan_device_len.argtypes = [vp, i32]
    L.spk_plan_device_len.restype = i64
    L.spk_plan_factor_phase.argtypes = [vp, i32]
    L.spk_plan_factor_phase.restype = i64
    L.spk_plan_solve_device.argtypes = [vp, vp, i64, i64, i32]
    L.spk_plan_solve_device.restype = i64
    L.spk_plan_solve_phase.argtypes = [vp, vp, i64, i64, i32]
    L.spk_plan_solve_phase.restype = i64
    L.spk_nccl_unique_id.argtypes = [vp]
    L.spk_nccl_unique_id.restype = i64
    L.spk_plan_comm_init.argtypes = [vp, vp]
    L.spk_plan_comm_init.restype = i64
    L.spk_plan_factor_multi.argtypes = [vp]
    L.spk_plan_factor_multi.restype = i64
    L.spk_plan_solve_multi.argtypes = [vp, vp, i64, i64]
    L.spk_plan_solve_multi.restype = i64
    L.spk_plan_condest.argtypes = [vp, vp, vp]
    L.spk_plan_condest.restype = dbl
    L.spk_cache_clear.restype = None
    L.spk_multi_create.argtypes = [i64, i64, I64P, I64P, I64P, I64P, I64P, vp, i32]
    L.spk_multi_create.restype = vp
    L.spk_multi_destroy.argtypes = [vp]
    L.spk_multi_destroy.restype = None
    L.spk_multi_plan.argtypes = [vp, i32]
    L.spk_multi_plan.restype = vp
    L.spk_multi_inmatrix.argtypes = [vp, i64, vp, F64P]
    L.spk_multi_inmatrix.restype = i64
    L.spk_multi_set_values.argtypes = [vp, F64P, vp]
    L.spk_multi_set_values.restype = i64
    L.spk_multi_factor.argtypes = [vp]
    L.spk_multi_factor.restype = i64
    L.spk_multi_get_factors.argtypes = [vp, vp, vp, vp]
    L.spk_multi_get_factors.restype = i64
    L.spk_multi_set_perm.argtypes = [vp, I64P, I64P]
    L.spk_multi_set_perm.restype = i64
    L.spk_multi_triangularsolve.argtypes = [vp, F64P, i64, i64]
    L.spk_multi_triangularsolve.restype = i64
    L.spk_plan_xchg_info.argtypes = [vp, i32, i64, vp]
    L.spk_plan_xchg_info.restype = i64
    L.spk_plan_stat.argtypes = [vp, i32]
    L.spk_plan_stat.restype = i64
    L.spk_plan_statf.argtypes = [vp, i32]
    L.spk_plan_statf.restype = dbl
    L.spk_last_error.restype = C.c_char_p
    L.spk_device_count.restype = i32
    L.spk_version.restype = C.c_char_p
    _lib = L
    return L


def last_error():
    return lib().spk_last_error().decode()


def _ptr(a):
    return None if a is None else a.ctypes.data


class Plan:
    """`spk_plan`: structure + factors resident on one GPU.  `base` is a `_SparseBase`-like object
    (n, nsuper, xsuper, snode, xlindx, lindx, xlnz, xunz, spd)."""

    def __init__(self, base, device=0, host_only=False, part=0, nparts=1):
        L = lib()
        self.L = L
        self.spd = bool(base.spd)
        self.n = int(base.n)
        self.nlnz = int(base.xlnz[base.n]) - 1
        self.nunz = 0 if self.spd else int(base.xunz[base.n]) - 1
        xunz = None if self.spd else base.xunz.ctypes.data
        self.h = L.spk_plan_create(base.n, base.nsuper, base.xsuper, base.snode, base.xlindx, base.lindx,
                                   base.xlnz, xunz, -1 if host_only else device, part, nparts)
        if not self.h:
            raise SpkError("spk_plan_create failed: " + last_error())
        self._perm_set = False

    def _ck(self, rc, what):
        if rc <= -100:
            raise SpkError(f"{what} failed ({rc}): {last_error()}")
        return rc

    def destroy(self):
        if getattr(self, "h", None):
            self.L.spk_plan_destroy(self.h)
            self.h = None

    __del__ = destroy

    def inmatrix(self, nzval, dest=None):
        nzval = np.ascontiguousarray(nzval, dtype=np.float64)
        if dest is not None:
            dest = np.ascontiguousarray(dest, dtype=np.int64)
        self._ck(self.L.spk_plan_inmatrix(self.h, nzval.size, _ptr(dest), nzval), "spk_plan_inmatrix")

    def set_matrix(self, A):
        """A: scipy.sparse matrix in the original ordering (kept on the device as CSR for residual / refine)."""
        A = A.tocsc(); A.sort_indices()
        colptr = np.ascontiguousarray(A.indptr, dtype=np.int64) + 1
        rowval = np.ascontiguousarray(A.indices, dtype=np.int64) + 1
        self._ck(self.L.spk_plan_set_matrix(self.h, A.nnz, colptr, rowval, np.ascontiguousarray(A.data, dtype=np.float64)), "spk_plan_set_matrix")

    def residual(self, b, x):
        """(res, relnorm): res = b - A x on the device, relnorm[q] = ||res_q|| / ||b_q||."""
        b = np.asfortranarray(b, dtype=np.float64); x = np.asfortranarray(x, dtype=np.float64)
        nrhs = 1 if b.ndim == 1 else b.shape[1]
        res = np.zeros_like(b, order="F"); rel = np.zeros(nrhs)
        self._ck(self.L.spk_plan_residual(self.h, b.ctypes.data, x.ctypes.data, nrhs, b.shape[0], res.ctypes.data, rel.ctypes.data), "spk_plan_residual")
        return res, rel

    def refine(self, b, x, maxit=3, tol=1e-15):
        """Iterative refinement of x in place with the resident factors; returns (corrections, relnorm)."""
        assert x.flags.f_contiguous or x.ndim == 1
        b = np.asfortranarray(b, dtype=np.float64)
        nrhs = 1 if b.ndim == 1 else b.shape[1]
        rel = np.zeros(nrhs)
        rc = self.L.spk_plan_refine(self.h, b.ctypes.data, x.ctypes.data, nrhs, b.shape[0], maxit, tol, rel.ctypes.data)
        if rc < 0:
            self._ck(rc, "spk_plan_refine")
        return int(rc), rel

    def condest(self):
        """(cond_1 estimate, ||A||_1, estimate of ||inv(A)||_1, solves used, lower_bound_only)"""
        out = np.zeros(2); info = np.zeros(2, np.int32)
        c = float(self.L.spk_plan_condest(self.h, out.ctypes.data, info.ctypes.data))
        if c < 0:
            raise SpkError("spk_plan_condest failed: " + last_error())
        return c, float(out[0]), float(out[1]), int(info[0]), bool(info[1])

    def reassemble(self):
        self._ck(self.L.spk_plan_reassemble(self.h), "spk_plan_reassemble")

    def solve_device(self, d_ptr, nrhs, ld, which=0):
        self._ck(self.L.spk_plan_solve_device(self.h, d_ptr, nrhs, ld, which), "spk_plan_solve_device")

    def set_values(self, lnz, unz=None):
        self._ck(self.L.spk_plan_set_values(self.h, lnz, _ptr(unz)), "spk_plan_set_values")

    def set_factors(self, lnz, unz=None, ipiv=None):
        self._ck(self.L.spk_plan_set_factors(self.h, lnz, _ptr(unz), _ptr(ipiv)), "spk_plan_set_factors")

    def factor(self):
        return int(self._ck(self.L.spk_plan_factor(self.h), "spk_plan_factor"))

    def get_factors(self, lnz=None, unz=None, ipiv=None):
        self._ck(self.L.spk_plan_get_factors(self.h, _ptr(lnz), _ptr(unz), _ptr(ipiv)), "spk_plan_get_factors")

    def solve(self, rhs, which=0):
        """rhs: (n,) or Fortran-ordered (n, nrhs) array in PERMUTED order; in place."""
        assert rhs.dtype == np.float64
        if rhs.ndim == 1:
            assert rhs.flags.c_contiguous
            nrhs, ld = 1, rhs.shape[0]
        else:
            assert rhs.flags.f_contiguous
            nrhs, ld = rhs.shape[1], rhs.shape[0]
        self._ck(self.L.spk_plan_solve(self.h, rhs.reshape(-1, order="A"), nrhs, ld, which), "spk_plan_solve")
        return rhs

    def set_perm(self, rperm, rinvp):
        self._ck(self.L.spk_plan_set_perm(self.h, np.ascontiguousarray(rperm, np.int64),
                                          np.ascontiguousarray(rinvp, np.int64)), "spk_plan_set_perm")
        self._perm_set = True

    def triangularsolve(self, b):
        """b: (n,) or Fortran-ordered (n, nrhs) in ORIGINAL order; in place."""
        assert b.dtype == np.float64
        if b.ndim == 1:
            assert b.flags.c_contiguous
            nrhs, ld = 1, b.shape[0]
        else:
            assert b.flags.f_contiguous
            nrhs, ld = b.shape[1], b.shape[0]
        self._ck(self.L.spk_plan_triangularsolve(self.h, b.reshape(-1, order="A"), nrhs, ld), "spk_plan_triangularsolve")
        return b

    # -- multi-part (one plan per GPU) -----------------------------------------------------
    def factor_phase(self, phase):
        return int(self._ck(self.L.spk_plan_factor_phase(self.h, phase), "spk_plan_factor_phase"))

    def solve_phase(self, d_ptr, nrhs, ld, phase):
        self._ck(self.L.spk_plan_solve_phase(self.h, d_ptr, nrhs, ld, phase), "spk_plan_solve_phase")

    def nccl_unique_id(self):
        ident = np.zeros(128, np.uint8)
        self._ck(self.L.spk_nccl_unique_id(ident.ctypes.data), "spk_nccl_unique_id")
        return ident

    def comm_init(self, ident):
        ident = np.ascontiguousarray(ident, dtype=np.uint8)
        self._ck(self.L.spk_plan_comm_init(self.h, ident.ctypes.data), "spk_plan_comm_init")

    def factor_multi(self):
        return int(self._ck(self.L.spk_plan_factor_multi(self.h), "spk_plan_factor_multi"))

    def solve_multi(self, d_ptr, nrhs, ld):
        self._ck(self.L.spk_plan_solve_multi(self.h, d_ptr, nrhs, ld), "spk_plan_solve_multi")

    def xchg_list(self, what):
        n = int(self.L.spk_plan_xchg_info(self.h, what, 0, None))
        out = []
        buf = np.zeros(8, np.int64)
        for i in range(n):
            self.L.spk_plan_xchg_info(self.h, what, i, buf.ctypes.data)
            out.append(buf.copy())
        return out

    def device_ptr(self, what):
        return self.L.spk_plan_device_ptr(self.h, what), int(self.L.spk_plan_device_len(self.h, what))

    def stat(self, what):
        return int(self.L.spk_plan_stat(self.h, what))

    def statf(self, what):
        return float(self.L.spk_plan_statf(self.h, what))


class MultiPlan:
    """`spk_multi`: one process, `ngpus` GPUs (devices 0..ngpus-1); same verbs as `Plan`."""

    def __init__(self, base, ngpus):
        L = lib()
        self.L = L
        self.spd = bool(base.spd)
        self.n = int(base.n)
        self.ngpus = int(ngpus)
        xunz = None if self.spd else base.xunz.ctypes.data
        self.h = L.spk_multi_create(base.n, base.nsuper, base.xsuper, base.snode, base.xlindx, base.lindx, base.xlnz, xunz, ngpus)
        if not self.h:
            raise SpkError("spk_multi_create failed: " + last_error())

    def _ck(self, rc, what):
        if rc <= -100:
            raise SpkError(f"{what} failed ({rc}): {last_error()}")
        return rc

    def destroy(self):
        if getattr(self, "h", None):
            self.L.spk_multi_destroy(self.h)
            self.h = None

    __del__ = destroy

    def inmatrix(self, nzval, dest=None):
        nzval = np.ascontiguousarray(nzval, dtype=np.float64)
        if dest is not None:
            dest = np.ascontiguousarray(dest, dtype=np.int64)
        self._ck(self.L.spk_multi_inmatrix(self.h, nzval.size, _ptr(dest), nzval), "spk_multi_inmatrix")

    def set_values(self, lnz, unz=None):
        self._ck(self.L.spk_multi_set_values(self.h, lnz, _ptr(unz)), "spk_multi_set_values")

    def factor(self):
        return int(self._ck(self.L.spk_multi_factor(self.h), "spk_multi_factor"))

    def get_factors(self, lnz=None, unz=None, ipiv=None):
        self._ck(self.L.spk_multi_get_factors(self.h, _ptr(lnz), _ptr(unz), _ptr(ipiv)), "spk_multi_get_factors")

    def set_perm(self, rperm, rinvp):
        self._ck(self.L.spk_multi_set_perm(self.h, np.ascontiguousarray(rperm, np.int64), np.ascontiguousarray(rinvp, np.int64)), "spk_multi_set_perm")

    def triangularsolve(self, b):
        assert b.dtype == np.float64
        if b.ndim == 1:
            nrhs, ld = 1, b.shape[0]
        else:
            assert b.flags.f_contiguous
            nrhs, ld = b.shape[1], b.shape[0]
        self._ck(self.L.spk_multi_triangularsolve(self.h, b.reshape(-1, order="A"), nrhs, ld), "spk_multi_triangularsolve")
        return b

    def part_stat(self, r, what):
        return int(self.L.spk_plan_stat(self.L.spk_multi_plan(self.h, r), what))

    def part_statf(self, r, what):
        return float(self.L.spk_plan_statf(self.L.spk_multi_plan(self.h, r), what))
