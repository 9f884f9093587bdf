"""Pins the CPU oracle and the host structure pipeline against the reference's own golden
vectors and known-answer tests (test/test_sparse_method.jl, test/test_graph.jl,
test/test_structunsymm.jl, test/test_small.jl)."""
import numpy as np
import pytest
import scipy.sparse as sp

import sparspak_jl_b200 as spk
import oracle
from common import maketridiagproblem, prepare, oracle_factor, residual, M

import json
import os

# fixtures: tests/golden/reference_goldens.json (each entry cites the reference test file:line it was copied from)
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")))
FIG313_I = GOLD["graph_fig313"]["I"]
FIG313_J = GOLD["graph_fig313"]["J"]


def test_graph_goldens():
    # test/test_graph.jl:20-21 (George & Liu Fig. 3.1.3)
    p = spk.Problem(6, 6)
    spk.insparse(p, FIG313_I, FIG313_J, [1.0] * 18)
    g = spk.Graph(p)
    assert g.xadj.tolist() == GOLD["graph_fig313"]["xadj"]
    assert g.adj.tolist() == GOLD["graph_fig313"]["adj"]
    assert spk.isstructuresymmetric(g)


def test_graph_symmetrisation_golden():
    # test/test_graph.jl:80-97: element (5,3) missing -> unsymmetric; symmetrised graph equals the full one
    I = [1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5, 5, 6, 6, 6]
    J = [1, 2, 6, 1, 2, 3, 4, 2, 3, 5, 2, 4, 5, 6, 1, 5, 6]
    p = spk.Problem(6, 6)
    spk.insparse(p, I, J, [1.0] * 17)
    g = spk.Graph(p)
    assert not spk.isstructuresymmetric(g)
    spk.makestructuresymmetric(g)
    assert spk.isstructuresymmetric(g)
    assert g.xadj.tolist() == [1, 3, 6, 8, 9, 11, 13]
    assert g.adj.tolist() == [2, 6, 1, 3, 4, 2, 5, 2, 3, 6, 1, 5]


def test_mmd_ordering_golden():
    # test/test_sparse_method.jl:87-90
    s = spk.SparseSolver(maketridiagproblem(11))
    spk.findorder(s)
    o = s.slvr.order
    assert o.rperm.tolist() == GOLD["tridiag11_mmd"]["rperm"]
    assert o.rinvp.tolist() == GOLD["tridiag11_mmd"]["rinvp"]
    assert o.cperm.tolist() == o.rperm.tolist() and o.cinvp.tolist() == o.rinvp.tolist()


def test_symbolic_structure_golden():
    # test/test_sparse_method.jl:127-131
    s = spk.SparseSolver(maketridiagproblem(11))
    spk.findorder(s); spk.symbolicfactor(s)
    b = s.slvr
    G = GOLD["tridiag11_symbolic"]
    assert b.xlnz.tolist() == G["xlnz"]
    assert b.xunz.tolist() == G["xunz"]
    assert b.xlindx.tolist() == G["xlindx"]
    assert b.lindx.tolist() == G["lindx"]


def test_inmatrix_golden():
    # test/test_sparse_method.jl:172-175
    s = spk.SparseSolver(maketridiagproblem(11))
    spk.findorder(s); spk.symbolicfactor(s); spk.inmatrix(s)
    assert s.slvr.unz.tolist() == GOLD["tridiag11_inmatrix"]["unz"]
    assert s.slvr.lnz.tolist() == GOLD["tridiag11_inmatrix"]["lnz"]


GOLD_LNZ = GOLD["tridiag11_lufactor"]["lnz"]


def test_oracle_lufactor_golden():
    # test/test_sparse_method.jl:219-227 — the factored lnz / unz of the 11x11 tridiagonal
    s = spk.SparseSolver(maketridiagproblem(11))
    spk.findorder(s); spk.symbolicfactor(s); spk.inmatrix(s)
    lnz, unz, ipiv, fl = oracle_factor(s.slvr)
    assert fl == 0
    g = np.array(GOLD_LNZ)
    assert np.linalg.norm(lnz - g) / np.linalg.norm(g) < 1e-15
    assert unz.tolist() == GOLD["tridiag11_lufactor"]["unz"]
    assert ipiv.tolist() == GOLD["tridiag11_lufactor"]["ipiv"]          # block-local LAPACK pivots (SURVEY.md §8a')


def test_oracle_worked_example_3x3_grid():
    # SURVEY.md §8a' worked example (emulated reference run): 3x3 5-point Laplacian, LU, MMD
    s = prepare(M.laplacian2d(3), False)
    b = s.slvr
    assert b.order.rperm.tolist() == [9, 7, 8, 3, 1, 2, 6, 4, 5]
    assert b.xsuper.tolist() == [1, 2, 3, 4, 5, 6, 10]
    assert b.xlindx.tolist() == [1, 4, 7, 11, 14, 17, 21]
    assert b.lindx.tolist() == [1, 3, 7, 2, 3, 8, 3, 7, 8, 9, 4, 6, 7, 5, 6, 8, 6, 7, 8, 9]
    assert b.xlnz.tolist() == [1, 4, 7, 11, 14, 17, 21, 25, 29, 33]
    assert b.xunz.tolist() == [1, 3, 5, 8, 10, 12, 12, 12, 12, 12]
    lnz, unz, ipiv, fl = oracle_factor(b)
    assert np.allclose(lnz[:3], [4, -0.25, -0.25]) and np.allclose(unz[:2], [-1, -1])
    assert np.allclose(lnz[6:10], [3.5, -0.0714, -0.0714, -0.2857], atol=1e-4)
    blk = lnz[16:32].reshape(4, 4, order="F")
    assert np.allclose(blk, [[3.5, -0.25, -0.25, -1], [-0.0714, 3.4643, -0.0357, -1.1429],
                             [-0.0714, -0.0103, 3.4639, -1.1546], [-0.2857, -0.3299, -0.3333, 2.6667]], atol=1e-4)
    assert ipiv[5:].tolist() == [1, 2, 3, 4]


@pytest.mark.parametrize("n", [11, 1101, 11000])
def test_oracle_tridiag_solves(n):
    # test/test_sparse_method.jl:264-376: solve vs the direct solution, tol 1e-6
    p = maketridiagproblem(n)
    s = prepare(p, False)
    lnz, unz, ipiv, fl = oracle_factor(s.slvr)
    x = oracle.triangularsolve(s.slvr, lnz, unz, ipiv, p.rhs.copy())
    import scipy.sparse.linalg as spla
    xr = spla.spsolve(p.csc().tocsc(), p.rhs)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-6
    assert residual(p.csc(), x, p.rhs) < 1e-14


def test_oracle_gridproblem():
    # test/test_sparse_method.jl:651-699: makegridproblem(5,3) (lower triangle only -> structurally unsymmetric)
    p = spk.makegridproblem(5, 3)
    a = p.csc()
    assert a.shape == (15, 15) and np.allclose(a.diagonal(), 8.0) and (sp.triu(a, 1).nnz == 0)
    s = prepare(p, False)
    lnz, unz, ipiv, fl = oracle_factor(s.slvr)
    b = a @ np.arange(1.0, 16.0)
    x = oracle.triangularsolve(s.slvr, lnz, unz, ipiv, b)
    assert np.allclose(x, np.arange(1.0, 16.0), rtol=1e-12)


def test_oracle_pivoting_fuzz():
    # test/test_structunsymm.jl:60-90 — random 4x4 sprand(4,4,0.3)+I systems that need real pivoting
    rng = np.random.default_rng(9876)
    nswaps = 0
    for k in range(300):
        a = sp.random(4, 4, density=0.3, random_state=rng, format="csc", data_rvs=rng.random) + sp.identity(4, format="csc")
        a = sp.csc_matrix(a); a.eliminate_zeros()
        if abs(np.linalg.det(a.toarray())) < 1e-8:
            continue
        s = prepare(a, False)
        lnz, unz, ipiv, fl = oracle_factor(s.slvr)
        b = rng.random(4)
        x = oracle.triangularsolve(s.slvr, lnz, unz, ipiv, b)
        assert np.linalg.norm(x - np.linalg.solve(a.toarray(), b)) < 1e-9
        loc = ipiv - (np.arange(4) - (s.slvr.xsuper[s.slvr.snode - 1] - 1))
        nswaps += int((loc != 1).sum())
    assert nswaps > 0            # the fuzz really exercises row interchanges


def test_oracle_structurally_unsymmetric_regression():
    # test/test_structunsymm.jl:13-57 style: unsymmetric pattern, explicit zeros appear in lnz/unz
    a = sp.csc_matrix(np.array([[2.0, 0, 0, 1], [1, 3, 0, 0], [0, 0, 4, 0], [0, 1, 2, 5]]))
    s = prepare(a, False)
    lnz, unz, ipiv, fl = oracle_factor(s.slvr)
    b = np.array([1.0, 2, 3, 4])
    x = oracle.triangularsolve(s.slvr, lnz, unz, ipiv, b)
    assert np.allclose(a @ x, b, atol=1e-14)


def test_oracle_ldlt_equals_lu_on_spd():
    # SPD contract: LDL^T L-blocks == LU L-blocks, D == diag(U), on SPD inputs
    for A, order in [(M.laplacian2d(12), None), (M.laplacian3d(8), spk.nd_grid_order(8, 8, 8))]:
        sl = prepare(A, False, order, maxblocksize=60)
        ss = prepare(A, True, order, maxblocksize=60)
        bl, bs = sl.slvr, ss.slvr
        assert np.array_equal(bl.xsuper, bs.xsuper) and np.array_equal(bl.lindx, bs.lindx) and np.array_equal(bl.xlnz, bs.xlnz)
        l_lu, u_lu, ipiv, _ = oracle_factor(bl)
        l_sp, _, _, fl = oracle_factor(bs)
        assert fl == 0 and np.all(ipiv == (np.arange(bl.n) - (bl.xsuper[bl.snode - 1] - 1) + 1))
        from common import spd_mask
        mask = spd_mask(bs)[: l_lu.size]
        assert np.abs(l_lu[mask] - l_sp[: l_lu.size][mask]).max() < 1e-13


def test_oracle_ldlt_residual():
    A = M.laplacian3d(9)
    s = prepare(A, True, spk.nd_grid_order(9, 9, 9))
    lnz, _, _, fl = oracle_factor(s.slvr)
    b = M.rhs_for(A)
    x = oracle.triangularsolve(s.slvr, lnz, None, None, b)
    assert residual(A, x, b) < 1e-14


def test_oracle_openblas_variant_matches_generic():
    # the CPU-baseline variant (dense call sites -> OpenBLAS, as Julia does for Float64) agrees with the generic loops
    A = M.convdiff3d(8)
    s = prepare(A, False, spk.nd_grid_order(8, 8, 8))
    l0, u0, p0, _ = oracle_factor(s.slvr)
    oracle.use_openblas(True)
    try:
        l1, u1, p1, _ = oracle_factor(s.slvr)
    finally:
        oracle.use_openblas(False)
    assert np.array_equal(p0, p1)
    assert np.abs(l0 - l1).max() < 1e-12 and np.abs(u0 - u1).max() < 1e-12
