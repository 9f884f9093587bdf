"""Shared helpers of the test-suite: problem builders, the oracle runs and comparisons."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

import sparspak_jl_b200 as spk
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I64P = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
F64P = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

FACTOR_RTOL = 1e-11       # north_star: factor entries agree to a relative 1e-11
RESID_TOL = 1e-12         # north_star: ||Ax-b|| / ||b|| <= 1e-12


def maketridiagproblem(n):
    """test/test_sparse_method.jl:66-78"""
    p = spk.Problem(n, n)
    for i in range(1, n):
        spk.inaij(p, i + 1, i, -1.0); spk.inaij(p, i, i, 4.0); spk.inaij(p, i, i + 1, -1.0); spk.inbi(p, i, 2.0 * i)
    spk.inaij(p, n, n, 4.0); spk.inbi(p, n, 3.0 * n + 1.0)
    return p


def prepare(A, spd, order=None, maxblocksize=None):
    """findorder! / symbolicfactor! / inmatrix! on the host; returns the solver."""
    s = (spk.SparseSpdSolver if spd else spk.SparseSolver)(A)
    if maxblocksize:
        s.slvr.maxblocksize = maxblocksize
    spk.findorder(s, order) if order is not None else spk.findorder(s)
    spk.symbolicfactor(s)
    spk.inmatrix(s)
    return s


def oracle_factor(b):
    """Oracle factors of the assembled values in b (a _SparseBase / _SparseSpdBase); returns (lnz, unz, ipiv, iflag)."""
    lnz = b.lnz.copy(); unz = b.unz.copy()
    if b.spd:
        fl = oracle.ldltfactor(b, lnz); ipiv = None
    else:
        ipiv = np.zeros(b.n, np.int64); fl = oracle.lufactor(b, lnz, unz, ipiv)
    return lnz, unz, ipiv, fl


def spd_mask(b):
    """True where an lnz entry is meaningful: everything except the strict upper triangle of the
    diagonal blocks of an LDL^T factor (don't-care in the reference)."""
    mask = np.ones(b.lnz.size, bool)
    if not b.spd:
        return mask
    mask[int(b.xlnz[b.n]) - 1:] = False          # the one surplus element of the SPD struct
    for k in range(b.nsuper):
        fj = b.xsuper[k] - 1; nj = b.xsuper[k + 1] - b.xsuper[k]
        jl = b.xlnz[fj + 1] - b.xlnz[fj]; base = b.xlnz[fj] - 1
        for c in range(1, nj):
            mask[base + c * jl: base + c * jl + c] = False
    return mask


def rel_err(a, ref, mask=None):
    if a.size == 0:
        return 0.0
    if mask is not None:
        a, ref = a[mask], ref[mask]
    scale = np.abs(ref).max()
    return float(np.abs(a - ref).max() / (scale if scale > 0 else 1.0))


def residual(A, x, b):
    return float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))


class HostSim:
    """ctypes wrapper of tests/hostsim (host executor of the device schedule)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
            L.sim_create.restype = C.c_void_p
            L.sim_create.argtypes = [C.c_int64, C.c_int64, I64P, I64P, I64P, I64P, I64P, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
            L.sim_factor.argtypes = [C.c_void_p, F64P, C.c_void_p, C.c_void_p]; L.sim_factor.restype = C.c_int64
            L.sim_solve.argtypes = [C.c_void_p, F64P]
            L.sim_stat.argtypes = [C.c_void_p, C.c_int]; L.sim_stat.restype = C.c_int64
            L.sim_statf.argtypes = [C.c_void_p, C.c_int]; L.sim_statf.restype = C.c_double
            L.sim_destroy.argtypes = [C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, b, relax_abs=4, relax_frac=0.02, dmma_buckets=True, alloc=True):
        self.b = b
        L = self.lib()
        self.h = L.sim_create(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz,
                              None if b.spd else b.xunz.ctypes.data, int(dmma_buckets), relax_abs, relax_frac, int(alloc))
        assert self.h

    def factor(self):
        b = self.b
        lnz = b.lnz[: int(b.xlnz[b.n]) - 1].copy(); unz = b.unz.copy(); ipiv = np.zeros(b.n, np.int64)
        fl = self.lib().sim_factor(self.h, lnz, None if b.spd else unz.ctypes.data, ipiv.ctypes.data)
        return lnz, unz, ipiv, int(fl)

    def solve(self, rhs):
        self.lib().sim_solve(self.h, rhs)
        return rhs

    def stat(self, k):
        return int(self.lib().sim_stat(self.h, k))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().sim_destroy(self.h); self.h = None


M = spk.matrices

# (name, matrix builder, spd?, ordering builder or None (= MMD), maxblocksize or None)
CASES = [
    ("tridiag11-lu", lambda: maketridiagproblem(11).csc(), False, None, None),
    ("lap2d-3-lu", lambda: M.laplacian2d(3), False, None, None),
    ("lap2d-3-spd", lambda: M.laplacian2d(3), True, None, None),
    ("lap2d-20-lu-mmd", lambda: M.laplacian2d(20), False, None, None),
    ("lap2d-20-spd-mmd", lambda: M.laplacian2d(20), True, None, None),
    ("lap2d-30-spd-nd", lambda: M.laplacian2d(30), True, lambda: spk.nd_grid_order(30, 30), None),
    ("lap3d-10-spd-nd", lambda: M.laplacian3d(10), True, lambda: spk.nd_grid_order(10, 10, 10), None),
    ("lap3d-10-lu-nd", lambda: M.laplacian3d(10), False, lambda: spk.nd_grid_order(10, 10, 10), None),
    ("convdiff-12-lu-nd", lambda: M.convdiff3d(12), False, lambda: spk.nd_grid_order(12, 12, 12), None),
    ("lap3d-14-spd-blk8", lambda: M.laplacian3d(14), True, lambda: spk.nd_grid_order(14, 14, 14), 8),
    ("convdiff-12-lu-blk5", lambda: M.convdiff3d(12), False, lambda: spk.nd_grid_order(12, 12, 12), 5),
    ("elasticity-5-spd", lambda: M.elasticity27(5), True, lambda: spk.nd_grid_order(5, 5, 5, 3), None),
    ("pivot-60-s0", lambda: M.pivoting_stress(60, 0.08, 0), False, None, None),
    ("pivot-60-s1", lambda: M.pivoting_stress(60, 0.08, 1), False, None, None),
    ("pivot-200-blk4", lambda: M.pivoting_stress(200, 0.03, 7), False, None, 4),
    ("pivot-4x4", lambda: M.pivoting_stress(4, 0.3, 3), False, None, None),
    ("single-1x1", lambda: sp.csc_matrix(np.array([[3.0]])), False, None, None),
    ("diag-5-spd", lambda: sp.diags([np.arange(1.0, 6.0)], [0], format="csc"), True, None, None),
    # one dense front: a single fundamental supernode split into several chunks (widths around maxblocksize)
    ("dense-70-lu", lambda: sp.csc_matrix(np.random.default_rng(5).standard_normal((70, 70)) + 0.5 * np.eye(70)), False, None, None),
    ("dense-130-spd", lambda: sp.csc_matrix((lambda G: G @ G.T + 130.0 * np.eye(130))(np.random.default_rng(6).standard_normal((130, 130)))), True, None, None),
]
