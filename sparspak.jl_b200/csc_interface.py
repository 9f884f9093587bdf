"""Host mirror of src/SparseCSCInterface/SparseCSCInterface.jl: `sparspaklu`,
`sparspaklu!` (here `sparspaklu_`), `ldiv!` (`ldiv`), and `\\` (`backslash`)."""
import numpy as np
import scipy.sparse as sp

from .sparse_solver import SparseSolver, findorder, symbolicfactor, inmatrix, factor, solve, _fail


def _csc(m):
    m = sp.csc_matrix(m); m.sort_indices()
    return m


def sparspaklu(m, factorize=True):
    """`sparspaklu(m; factorize=true)` (SparseCSCInterface.jl:218-227)."""
    lu = SparseSolver(_csc(m))
    if factorize:
        findorder(lu) or _fail("Finding Order.")
        symbolicfactor(lu) or _fail("Symbolic Factorization.")
        inmatrix(lu) or _fail("Matrix input.")
        factor(lu) or _fail("Numerical Factorization.")
    return lu


def sparspaklu_(lu, m, allow_pattern_change=True):
    """`sparspaklu!(lu, m; allow_pattern_change=true)` (SparseCSCInterface.jl:247-268): reuse
    ordering + symbolic factorisation when the pattern is unchanged."""
    m = _csc(m)
    old = lu.p
    changed = (not sp.issparse(old) or old.shape != m.shape or
               not np.array_equal(m.indptr, old.indptr) or not np.array_equal(m.indices, old.indices))
    if changed:
        if allow_pattern_change or not lu._symbolicdone:
            fresh = SparseSolver(m)
            lu.slvr._destroy_plan()
            lu.__dict__.update(fresh.__dict__)
        else:
            raise RuntimeError("'allow_pattern_change=false', but sparsity pattern of matrix 'm' "
                               "does not match that used to create 'lu'")
    lu.p = m
    lu._orderdone or findorder(lu) or _fail("Finding Order.")
    lu._symbolicdone or symbolicfactor(lu) or _fail("Symbolic Factorization.")
    lu._inmatrixdone = False
    lu._factordone = False
    lu._trisolvedone = False
    inmatrix(lu) or _fail("Matrix input.")
    factor(lu) or _fail("Numerical Factorization.")
    return lu


def ldiv(*args):
    """`ldiv!(u, lu, v)` / `ldiv!(lu, v)` (SparseCSCInterface.jl:277-293)."""
    if len(args) == 3:
        u, lu, v = args
        u[:] = v
    else:
        lu, u = args
    solve(lu, u)
    lu._trisolvedone = False
    return u


def backslash(lu, v):
    """`lu \\ v` (SparseCSCInterface.jl:300)."""
    u = np.array(v, dtype=np.float64)
    return ldiv(lu, u)
