"""Solver data + step implementations: host mirror of `_SparseBase` (LU,
src/SparseMethod/SpkSparseBase.jl) and `_SparseSpdBase` (LDL^T,
src/SparseSpdMethod/SpkSparseSpdBase.jl).

`_findorder`, `_symbolicfactor`, `_inmatrix` are host code here as in the reference.
`_factor` and `_triangularsolve` are THE BOUNDARY: they call the CUDA
C-ABI (include/spk_b200.h) through a plan that keeps structure and factors resident in
HBM.  There is no CPU fallback: without the CUDA library / a GPU they raise.
"""
import numpy as np
import scipy.sparse as sp

from . import _hostlib
from .problem import Problem, Graph, Ordering, ETree, makestructuresymmetric, mmd


def _csc_1based(m):
    m = sp.csc_matrix(m)
    m.sum_duplicates()               # one entry per (i, j), as SparseMatrixCSC guarantees
    m.sort_indices()
    return m.indptr.astype(np.int64) + 1, m.indices.astype(np.int64) + 1, np.ascontiguousarray(m.data, dtype=np.float64)


class _Base:
    spd = False

    def __init__(self, p, maxblocksize):
        self.maxblocksize = int(maxblocksize)       # "can be set by the user"
        self.n = p.nrows if isinstance(p, Problem) else p.shape[1]
        self.nnz = p.nnz
        self.nnzl = 0
        self.nsub = 0
        self.nsuper = 1 if self.n > 0 else 0
        self.errflag = 0
        self.order = Ordering(self.n)
        self.g = Graph(p)
        self.t = ETree(self.n)
        e = np.zeros(0, np.int64)
        self.colcnt = e; self.snode = e; self.xsuper = e; self.xlindx = e; self.lindx = e
        self.xlnz = e; self.xunz = e; self.ipiv = e
        self.lnz = np.zeros(0); self.unz = np.zeros(0)
        self._plan = None           # device plan (rebuilt by _symbolicfactor)
        self._dest = None           # inmatrix index map, built once per pattern
        self._dest_key = None
        self._factors_on_host = False

    # -- step 1 -------------------------------------------------------------
    def _findorder(self, orderfunction=mmd):
        """`_findorder!` (SpkSparseBase.jl:173-185)."""
        if self.n == 0:
            raise RuntimeError("An empty problem, no ordering found.")
        makestructuresymmetric(self.g)
        if callable(orderfunction):
            orderfunction(self.g, self.order)
        else:                        # findorderperm!: an explicit 1-based permutation
            perm = np.asarray(orderfunction, np.int64)
            self.order.rperm[:] = perm
            self.order.rinvp[perm - 1] = np.arange(1, self.n + 1)
            self.order.cperm[:] = self.order.rperm; self.order.cinvp[:] = self.order.rinvp
        return True

    # -- step 2 -------------------------------------------------------------
    def _symbolicfactor(self):
        """`_symbolicfactor!` (SpkSparseBase.jl:193-251 / SpkSparseSpdBase.jl:178-232)."""
        if self.n == 0:
            raise RuntimeError("An empty problem. No symbolic factorization done.")
        H = _hostlib.lib()
        n, g, o, t = self.n, self.g, self.order, self.t
        self.colcnt = np.zeros(n, np.int64)
        self.snode = np.zeros(n, np.int64)
        xsuper = np.zeros(n + 1, np.int64)
        H.spkh_etree(n, g.xadj, g.adj, o.rperm, o.rinvp, t.parent)
        H.spkh_postorder(n, t.parent, o.rperm, o.rinvp, None)
        self.nnzl = int(H.spkh_colcounts(n, g.xadj, g.adj, o.rperm, o.rinvp, t.parent, self.colcnt))
        H.spkh_postorder(n, t.parent, o.rperm, o.rinvp, self.colcnt.ctypes.data)
        o.cperm = o.rperm; o.cinvp = o.rinvp
        out = np.zeros(2, np.int64)
        H.spkh_findsupernodes(n, t.parent, self.colcnt, self.maxblocksize, xsuper, self.snode, out)
        self.nsuper, self.nsub = int(out[0]), int(out[1])
        self.xsuper = xsuper[: self.nsuper + 1].copy()
        self.lindx = np.zeros(self.nsub, np.int64)
        self.xlindx = np.zeros(self.nsuper + 1, np.int64)
        self.xlnz = np.zeros(n + 1, np.int64)
        if self.spd:
            H.spkh_nonzeroindexs(n, self.colcnt, self.nsuper, self.xsuper, self.xlnz, None)
        else:
            self.xunz = np.zeros(n + 1, np.int64)
            self.ipiv = np.zeros(n, np.int64)
            H.spkh_nonzeroindexs(n, self.colcnt, self.nsuper, self.xsuper, self.xlnz, self.xunz.ctypes.data)
        rc = H.spkh_symbolicfact(n, g.xadj, g.adj, o.rperm, o.rinvp, self.colcnt, self.nsuper, self.xsuper,
                                 self.snode, self.xlindx, self.lindx)
        if rc != 0:
            raise RuntimeError("Inconsistency in data structure.")
        # the SPD struct allocates one element more than it uses (SpkSparseSpdBase.jl:226)
        self.lnz = np.zeros(int(self.xlnz[n]) - 1 + (1 if self.spd else 0))
        self.unz = np.zeros(0 if self.spd else int(self.xunz[n]) - 1)
        self._destroy_plan()
        self._dest = None
        return True

    # -- step 3 -------------------------------------------------------------
    def _inmatrix_map(self, p):
        colptr, rowval, nzval = _csc_1based(p.csc() if isinstance(p, Problem) else p)
        key = (rowval.size, hash(colptr.tobytes()), hash(rowval.tobytes()))     # the map belongs to ONE pattern
        if self._dest is None or self._dest_key != key:
            H = _hostlib.lib()
            dest = np.zeros(rowval.size, np.int64)
            o = self.order
            if self.spd:
                bad = H.spkh_inmatrix_map_spd(self.n, colptr, rowval, o.rinvp, o.cinvp, self.snode, self.xsuper,
                                              self.xlindx, self.lindx, self.xlnz, dest)
            else:
                bad = H.spkh_inmatrix_map_lu(self.n, colptr, rowval, o.rinvp, o.cinvp, self.snode, self.xsuper,
                                             self.xlindx, self.lindx, self.xlnz, self.xunz, dest)
            if bad:
                i = int(rowval[bad - 1]); j = int(np.searchsorted(colptr, bad, side="right"))
                raise RuntimeError(f"No space for matrix element ({o.rinvp[i - 1]}, {o.cinvp[j - 1]}).")
            self._dest = dest; self._dest_key = key
        return self._dest, nzval

    def _inmatrix(self, p):
        """`_inmatrix!` (SpkSparseBase.jl:302-372, SparseCSCInterface.jl:101-169,
        SpkSparseSpdBase.jl:234-311): scatter A into the rectangular supernode storage."""
        if self.n == 0:
            raise RuntimeError("An empty problem. No matrix.")
        dest, nzval = self._inmatrix_map(p)
        self.lnz[:] = 0.0
        self.unz[:] = 0.0
        if not self.spd:
            self.ipiv[:] = 0
        _hostlib.lib().spkh_scatter_values(nzval.size, dest, nzval, self.lnz,
                                           None if self.spd else self.unz.ctypes.data)
        self._nzval = nzval
        self._factors_on_host = False
        return True

    # -- steps 4, 5: the boundary --------------------------------------------
    def _get_plan(self):
        from . import _cudalib
        if self._plan is None:
            self._plan = _cudalib.Plan(self)
        return self._plan

    def _destroy_plan(self):
        if self._plan is not None:
            self._plan.destroy()
            self._plan = None

    def _factor(self):
        """`_factor!` (SpkSparseBase.jl:378-391 / SpkSparseSpdBase.jl:313-332) -> CUDA."""
        if self.n == 0:
            raise RuntimeError("An empty problem. No matrix.")
        plan = self._get_plan()
        plan.set_values(self.lnz, None if self.spd else self.unz)
        self.errflag = plan.factor()
        # in-place overwrite of lnz/unz/ipiv is the reference's contract
        plan.get_factors(self.lnz, None if self.spd else self.unz, None if self.spd else self.ipiv)
        self._factors_on_host = True
        if self.errflag != 0:
            raise RuntimeError("An empty problem. No matrix.")     # sic (SpkSparseBase.jl:387)
        return True

    def _triangularsolve(self, solution):
        """`_triangularsolve!` (SpkSparseBase.jl:400-416 / SpkSparseSpdBase.jl:334-356)."""
        if self.n == 0:
            raise RuntimeError("An empty problem. No solution.")
        plan = self._get_plan()
        plan.set_perm(self.order.rperm, self.order.rinvp)
        plan.triangularsolve(solution)
        return True

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass


class _SparseBase(_Base):
    """LU (SpkSparseBase.jl:99-171; maxblocksize 30 at :133)."""
    spd = False

    def __init__(self, p):
        super().__init__(p, 30)


class _SparseSpdBase(_Base):
    """LDL^T (SpkSparseSpdBase.jl:84-139; maxblocksize 60 at :111)."""
    spd = True

    def __init__(self, p):
        super().__init__(p, 60)
