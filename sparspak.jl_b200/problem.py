"""Host mirror of the reference's input data model: `Problem`, `Graph`, `Ordering`,
`ETree`, `Grid` (src/Problem/SpkProblem.jl, src/Graph/SpkGraph.jl,
src/Ordering/SpkOrdering.jl, src/ETree/SpkETree.jl, src/Grid/SpkGrid.jl).

These stay host code in the reference too (north_star); they exist here only so
the reference's own call sequence can be driven from Python.  Julia's `f!`
names lose the bang: `inaij!` -> `inaij`, `makerhs!` -> `makerhs`, ...
All index arrays are 1-based int64, exactly as the reference stores them.
"""
import numpy as np
import scipy.sparse as sp

from . import _hostlib


class Problem:
    """`Problem{IT,FT}` (SpkProblem.jl:105-170): column lists with sorted row
    subscripts; values of repeated (i,j) are ADDED (SpkProblem.jl:226-236);
    explicitly stored zeros count as structural entries."""

    def __init__(self, nrows=0, ncols=0, nnz=2500, z=0.0, info=""):
        self.info = info
        self.nrows = int(nrows)
        self.ncols = int(ncols)
        self.dtype = np.float64
        self._I, self._J, self._V = [], [], []        # pending coefficient chunks
        self._csc = None
        self.rhs = np.zeros(self.nrows)
        self.x = np.zeros(self.ncols)

    # -- coefficients ------------------------------------------------------
    def _push(self, I, J, V):
        self._I.append(np.asarray(I, dtype=np.int64).ravel())
        self._J.append(np.asarray(J, dtype=np.int64).ravel())
        self._V.append(np.asarray(V, dtype=np.float64).ravel())
        self._csc = None

    def csc(self):
        """Column-sorted matrix with duplicates summed and stored zeros kept."""
        if self._csc is None:
            if self._I:
                I = np.concatenate(self._I); J = np.concatenate(self._J); V = np.concatenate(self._V)
            else:
                I = J = np.zeros(0, np.int64); V = np.zeros(0)
            m = sp.coo_matrix((V, (I - 1, J - 1)), shape=(self.nrows, self.ncols)).tocsc()
            m.sum_duplicates(); m.sort_indices()
            self._csc = m
            self._I, self._J, self._V = [I], [J], [V]
        return self._csc

    @property
    def nnz(self):
        return int(self.csc().nnz)

    def _grow(self, nrows, ncols):
        if nrows > self.nrows:
            self.rhs = np.concatenate([self.rhs, np.zeros(nrows - self.nrows)]); self.nrows = nrows
        if ncols > self.ncols:
            self.x = np.concatenate([self.x, np.zeros(ncols - self.ncols)]); self.ncols = ncols


def inaij(p, rnum, cnum, aij=0.0):
    """`inaij!` (SpkProblem.jl:177-251): add a coefficient; invalid subscripts are ignored."""
    if rnum < 1 or cnum < 1:
        return False
    p._grow(int(rnum), int(cnum))
    p._push([rnum], [cnum], [aij])
    return True


def inbi(p, rnum, bi):
    """`inbi!` (SpkProblem.jl:258-271)."""
    if rnum < 1:
        raise ValueError(f"Invalid rhs subscript {rnum}.")
    p._grow(int(rnum), p.ncols)
    p.rhs[rnum - 1] += bi
    return True


def insparse(p, *args):
    """`insparse!` (SpkProblem.jl:280-298): from a scipy sparse matrix or 1-based (I, J, V)."""
    if len(args) == 1:
        m = sp.coo_matrix(args[0])
        I, J, V = m.row.astype(np.int64) + 1, m.col.astype(np.int64) + 1, m.data
    else:
        I, J, V = args
    I = np.asarray(I, np.int64); J = np.asarray(J, np.int64)
    if I.size and (I.min() < 1 or J.min() < 1):
        return False
    if I.size:
        p._grow(int(I.max()), int(J.max()))
    p._push(I, J, V)
    return True


def outsparse(p):
    """`outsparse` (SpkProblem.jl:305-326)."""
    return p.csc().copy()


def infullrhs(p, rhs):
    """`infullrhs!` (SpkProblem.jl:512-517)."""
    p.rhs = np.array(rhs, dtype=np.float64).copy()
    return True


def computeresidual(p, res, xin=None, mtype="T"):
    """`computeresidual` (SpkProblem.jl:448-496): res = rhs - A*x, where an entry (r,c), r != c,
    also acts as (c,r) — the reference's flag is 1 for every accepted `mtype`."""
    if mtype.lower() not in ("t", "l", "u"):
        raise ValueError(f"Invalid value for mtype, {mtype}.")
    # The reference always ends up using p.x here (its `isempty(xin)` test is inverted,
    # SpkProblem.jl:470-474); makerhs! stores x into p.x first, so the result is the same.
    x = p.x
    a = p.csc()
    off = a - sp.diags(a.diagonal(), format="csc")
    res[:] = p.rhs - a @ x - off.T @ x
    return True


def makerhs(p, x=None, mtype="T"):
    """`makerhs!` (SpkProblem.jl:402-423): rhs := A_sym * x with x = 1..n by default."""
    if p.nnz == 0:
        raise ValueError("Matrix is NULL. The rhs cannot be computed.")
    p.x = np.arange(1, p.ncols + 1, dtype=np.float64) if x is None or len(x) == 0 else np.array(x, np.float64)
    p.rhs = np.zeros(p.nrows)
    res = np.zeros(p.nrows)
    computeresidual(p, res, p.x, mtype)
    p.rhs = -res
    p.x = np.zeros(p.ncols)
    return p


class Grid:
    """`Grid` (SpkGrid.jl:19-28): v[i,j] = k*(i-1)+j."""

    def __init__(self, h, k):
        self.h, self.k = int(h), int(k)
        self.v = (np.arange(h)[:, None] * k + np.arange(1, k + 1)[None, :]).astype(np.int64)


def makegridproblem(h, k=None):
    """`makegridproblem` (SpkProblem.jl:340-379): lower triangle of a 9-point stencil, diagonal 8."""
    g = h if isinstance(h, Grid) else Grid(h, k)
    p = Problem(g.h * g.k, g.h * g.k)
    v = g.v
    I, J, V = [], [], []
    for i in range(g.h):
        for j in range(g.k):
            I.append(v[i, j]); J.append(v[i, j]); V.append(8.0)
            if i > 0: I.append(v[i, j]); J.append(v[i - 1, j]); V.append(-1.0)
            if j > 0: I.append(v[i, j]); J.append(v[i, j - 1]); V.append(-1.0)
            if i > 0 and j > 0: I.append(v[i, j]); J.append(v[i - 1, j - 1]); V.append(-1.0)
            if j < g.k - 1 and i > 0: I.append(v[i, j]); J.append(v[i - 1, j + 1]); V.append(-1.0)
    insparse(p, I, J, V)
    return p


class Graph:
    """`Graph{IT}` (SpkGraph.jl:31-87, CSC twin SparseCSCInterface.jl:12-56): adjacency
    lists of the matrix columns without the diagonal, 1-based."""

    def __init__(self, p, diagonal=False):
        a = p.csc() if isinstance(p, Problem) else sp.csc_matrix(p)
        a.sort_indices()
        self.nv = a.shape[1]
        self.nrows, self.ncols = a.shape
        colptr = a.indptr.astype(np.int64); rowval = a.indices.astype(np.int64)
        cols = np.repeat(np.arange(self.nv, dtype=np.int64), np.diff(colptr))
        keep = np.ones(rowval.size, bool) if diagonal else (rowval != cols)
        cnt = np.bincount(cols[keep], minlength=self.nv)
        self.xadj = np.concatenate([[1], 1 + np.cumsum(cnt)]).astype(np.int64)
        self.adj = (rowval[keep] + 1).astype(np.int64)
        self.nedges = int(self.adj.size)


def isstructuresymmetric(g):
    """`isstructuresymmetric` (SpkGraph.jl:298-316)."""
    n = g.nv
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(g.xadj))
    a = sp.csc_matrix((np.ones(g.adj.size, np.int8), (g.adj - 1, cols)), shape=(n, n))
    return (a != a.T).nnz == 0


def makestructuresymmetric(g):
    """`makestructuresymmetric` (SpkGraph.jl:94-268): pattern of A + A^T, lists sorted."""
    if isstructuresymmetric(g):
        return True
    n = g.nv
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(g.xadj))
    a = sp.csc_matrix((np.ones(g.adj.size, np.int8), (g.adj - 1, cols)), shape=(n, n))
    s = (a + a.T).tocsc(); s.sort_indices()
    g.xadj = (s.indptr.astype(np.int64) + 1)
    g.adj = (s.indices.astype(np.int64) + 1)
    g.nedges = int(g.adj.size)
    return True


class Ordering:
    """`Ordering{IT}` (SpkOrdering.jl:66-126): identity row/column permutations."""

    def __init__(self, nrows, ncols=None):
        ncols = nrows if ncols is None else ncols
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.rperm = np.arange(1, nrows + 1, dtype=np.int64)
        self.rinvp = self.rperm.copy()
        self.cperm = np.arange(1, ncols + 1, dtype=np.int64)
        self.cinvp = self.cperm.copy()


class ETree:
    """`ETree{IT}` (SpkETree.jl:20-33)."""

    def __init__(self, nv):
        self.nv = int(nv)
        self.parent = np.zeros(nv, dtype=np.int64)


def mmd(g, order):
    """`mmd!` (SpkMMD.jl:38-42): multiple minimum degree ordering of a symmetric graph."""
    _hostlib.lib().spkh_mmd(g.nv, g.xadj, g.adj, order.rperm, order.rinvp)
    order.cinvp[:] = order.rinvp
    order.cperm[:] = order.rperm


def nd_grid_order(nx, ny=1, nz=1, dof=1, leaf=8):
    """Geometric nested-dissection `orderfunction(g, order)` for an nx*ny*nz grid (x fastest,
    `dof` unknowns per node).  Not in the reference (it has MMD only); it plugs into the
    reference's ordering-callback seam `findorder!(s, orderfunction)` (SpkSparseSolver.jl:103-111)."""
    def orderfunction(g, order):
        n = nx * ny * nz * dof
        if n != g.nv:
            raise ValueError("grid size does not match the graph")
        rc = _hostlib.lib().spkh_nd_grid(nx, ny, nz, dof, leaf, order.rperm, order.rinvp)
        if rc != 0:
            raise RuntimeError("nested dissection failed")
        order.cinvp[:] = order.rinvp
        order.cperm[:] = order.rperm
    return orderfunction
