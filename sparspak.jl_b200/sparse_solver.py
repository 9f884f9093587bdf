"""Solver state machines: host mirror of `SparseSolver` (src/SparseMethod/SpkSparseSolver.jl)
and `SparseSpdSolver` (src/SparseSpdMethod/SpkSparseSpdSolver.jl) — done-flags and
sequencing errors only; same verbs without Julia's bang."""
import numpy as np
import scipy.sparse as sp

from .problem import Problem
from .sparse_base import _SparseBase, _SparseSpdBase


class SequenceError(RuntimeError):
    """The reference throws ErrorException("Sequence error. ...") (SpkSparseSolver.jl:167-243)."""


class _Solver:
    _base = None

    def __init__(self, p):
        if not isinstance(p, Problem) and not sp.issparse(p):
            raise TypeError("SparseSolver needs a Problem or a scipy sparse matrix")
        if sp.issparse(p):
            p = sp.csc_matrix(p); p.sort_indices()
        self.p = p
        self.slvr = self._base(p)
        self.n = self.slvr.n
        self._inmatrixdone = self._orderdone = self._symbolicdone = False
        self._factordone = self._trisolvedone = self._refinedone = self._condestdone = False


def findorder(s, orderfunction=None):
    """`findorder!` (SpkSparseSolver.jl:103-133)."""
    if s._orderdone:
        return True
    if orderfunction is None:
        s.slvr._findorder()
    else:
        s.slvr._findorder(orderfunction)
    s._orderdone = True
    s._symbolicdone = False
    return True


def findorderperm(s, perm):
    """`findorderperm!` (SpkSparseSolver.jl:144-152)."""
    return findorder(s, np.asarray(perm, np.int64))


def symbolicfactor(s):
    """`symbolicfactor!` (SpkSparseSolver.jl:163-175)."""
    if s._symbolicdone:
        return True
    if not s._orderdone:
        raise SequenceError("Sequence error. Ordering not done yet.")
    s.slvr._symbolicfactor()
    s._symbolicdone = True
    s._inmatrixdone = False
    return True


def inmatrix(s):
    """`inmatrix!` (SpkSparseSolver.jl:188-199)."""
    if s._inmatrixdone:
        return True
    if not s._symbolicdone:
        raise SequenceError("Sequence error. Symbolic factor not done yet.")
    ok = s.slvr._inmatrix(s.p)
    s._inmatrixdone = True
    s._factordone = False
    return ok


def factor(s):
    """`factor!` (SpkSparseSolver.jl:209-227)."""
    if s._factordone:
        return True
    if not s._inmatrixdone:
        raise SequenceError("Sequence error. Matrix input not done yet.")
    s._trisolvedone = False
    s.slvr._factor()
    if s.slvr.errflag == 0:
        s._factordone = True
        return True
    return False


def triangularsolve(s, rhs=None):
    """`triangularsolve!` (SpkSparseSolver.jl:237-275).  With `rhs` the solve always runs and
    overwrites `rhs`; without, the Problem's rhs is solved once into `p.x`."""
    if rhs is None:
        if s._trisolvedone:
            return True
        if not s._factordone:
            raise SequenceError("Sequence error. Factorization not done yet.")
        temp = np.array(s.p.rhs[: s.p.nrows], dtype=np.float64)
        assert temp.size == s.n
        triangularsolve(s, temp)
        s.p.x[:] = temp
        s._trisolvedone = True
        s._refinedone = False
        return True
    if not s._factordone:
        raise SequenceError("Sequence error. Factorization not done yet.")
    s.slvr._triangularsolve(rhs)
    s._trisolvedone = True
    s._refinedone = False
    return True


def solve(s, rhs=None):
    """`solve!` (SpkSparseSolver.jl:80-87; with rhs: SparseCSCInterface.jl:194-201)."""
    findorder(s) or _fail("Finding Order.")
    symbolicfactor(s) or _fail("Symbolic Factorization.")
    inmatrix(s) or _fail("Matrix input.")
    factor(s) or _fail("Numerical Factorization.")
    triangularsolve(s, rhs) or _fail("Triangular Solve.")
    return True


def _fail(msg):
    raise RuntimeError(msg)


class SparseSolver(_Solver):
    """LU general sparse solver (SpkSparseSolver.jl:16-62; CSC ctor SparseCSCInterface.jl:171-186)."""
    _base = _SparseBase


class SparseSpdSolver(_Solver):
    """LDL^T solver (SpkSparseSpdSolver.jl:21-62).  The reference accepts only a Problem
    (SpkSparseSpdSolver.jl:40); a scipy matrix is accepted here as a convenience."""
    _base = _SparseSpdBase
