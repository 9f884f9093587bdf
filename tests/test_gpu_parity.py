"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
inputs.  Bars (north_star): pivot sequence and all integer structure bit-exact, factor entries
to a relative 1e-11, solve residual ||Ax-b||/||b|| <= 1e-12."""
import numpy as np
import pytest
import scipy.sparse as sp

import sparspak_jl_b200 as spk
from sparspak_jl_b200 import _cudalib
import oracle
from common import (CASES, prepare, oracle_factor, spd_mask, rel_err, residual, M, maketridiagproblem,
                    FACTOR_RTOL, RESID_TOL)

pytestmark = pytest.mark.gpu


def gpu_factor_plan(b):
    plan = _cudalib.Plan(b)
    plan.set_values(b.lnz, None if b.spd else b.unz)
    fl = plan.factor()
    nl = int(b.xlnz[b.n]) - 1
    lnz = np.zeros(b.lnz.size); unz = np.zeros(b.unz.size); ipiv = np.zeros(b.n, np.int64)
    plan.get_factors(lnz, None if b.spd else unz, None if b.spd else ipiv)
    return plan, lnz, unz, ipiv, fl


@pytest.mark.parametrize("name,build,spd,order,maxblk", CASES, ids=[c[0] for c in CASES])
def test_factor_and_solve_match_oracle(name, build, spd, order, maxblk):
    A = build()
    s = prepare(A, spd, order() if order else None, maxblk)
    b = s.slvr
    lo, uo, po, fo = oracle_factor(b)
    plan, lg, ug, pg, fg = gpu_factor_plan(b)
    assert fg == fo == 0
    assert rel_err(lg, lo, spd_mask(b)) < FACTOR_RTOL
    if not spd:
        assert np.array_equal(pg, po), "pivot sequence differs from the reference rule"
        assert rel_err(ug, uo) < FACTOR_RTOL
    bb = M.rhs_for(A)
    x = bb.copy()
    plan.set_perm(b.order.rperm, b.order.rinvp)
    plan.triangularsolve(x)
    xo = oracle.triangularsolve(b, lo, uo, po, bb)
    assert residual(A, x, bb) < RESID_TOL
    assert np.linalg.norm(x - xo) / np.linalg.norm(xo) < 1e-10
    plan.destroy()


def test_golden_tridiagonal_factors_on_gpu():
    # test/test_sparse_method.jl:219-227 directly against the CUDA path
    from test_oracle_golden import GOLD_LNZ
    s = prepare(maketridiagproblem(11), False)
    plan, lg, ug, pg, fg = gpu_factor_plan(s.slvr)
    g = np.array(GOLD_LNZ)
    assert np.linalg.norm(lg - g) / np.linalg.norm(g) < 1e-14
    assert np.abs(ug + 1.0).max() == 0.0
    assert pg.tolist() == [1] * 10 + [2]


@pytest.mark.parametrize("spd", [False, True])
def test_stateless_dropins(spd):
    """spk_lufactor_f64 / spk_lulsolve_f64 / spk_luusolve_f64 / spk_ldltfactor_f64 / spk_ldltsolve_f64:
    the Julia signatures of _lufactor! etc. (SpkSparseBase.jl:384,409-411; SpkSparseSpdBase.jl:325,351)."""
    A = M.convdiff3d(9) if not spd else M.laplacian3d(9)
    s = prepare(A, spd, spk.nd_grid_order(9, 9, 9))
    b = s.slvr
    L = _cudalib.lib()
    lo, uo, po, _ = oracle_factor(b)
    lnz = b.lnz.copy(); unz = b.unz.copy(); ipiv = np.zeros(b.n, np.int64)
    bb = M.rhs_for(A)
    rhs = np.ascontiguousarray(bb[b.order.rperm - 1])
    if spd:
        assert L.spk_ldltfactor_f64(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, lnz) == 0
        assert rel_err(lnz, lo, spd_mask(b)) < FACTOR_RTOL
        assert L.spk_ldltsolve_f64(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, rhs) == 1
    else:
        assert L.spk_lufactor_f64(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, ipiv) == 0
        assert np.array_equal(ipiv, po) and rel_err(lnz, lo) < FACTOR_RTOL and rel_err(unz, uo) < FACTOR_RTOL
        # forward sweep alone == oracle's _lulsolve!
        fw = rhs.copy(); fo = rhs.copy()
        assert L.spk_lulsolve_f64(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, ipiv, fw) == 1
        oracle.lib().spko_lulsolve(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lo, po, fo)
        assert np.linalg.norm(fw - fo) / np.linalg.norm(fo) < 1e-12
        assert L.spk_luusolve_f64(b.n, b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, fw) == 1
        rhs = fw
    x = rhs[b.order.rinvp - 1]
    assert residual(A, x, bb) < RESID_TOL


def test_solver_api_dropin_lu():
    """The reference call sequence (findorder!/symbolicfactor!/inmatrix!/factor!/triangularsolve!/solve!)
    through the host mirror; lnz/unz/ipiv of the solver are overwritten in place as in the reference."""
    p = maketridiagproblem(1101)
    s = spk.SparseSolver(p)
    assert spk.solve(s)
    import scipy.sparse.linalg as spla
    xr = spla.spsolve(p.csc().tocsc(), p.rhs)
    assert np.linalg.norm(p.x - xr) / np.linalg.norm(xr) < 1e-6          # the reference's own bar
    assert residual(p.csc(), p.x, p.rhs) < RESID_TOL
    lo, uo, po, _ = None, None, None, None
    assert s._factordone and s._trisolvedone
    assert s.slvr.ipiv.min() >= 1


def test_solver_api_dropin_spd_and_nd_callback():
    A = M.laplacian2d(40)
    s = spk.SparseSpdSolver(A)
    spk.findorder(s, spk.nd_grid_order(40, 40))
    spk.symbolicfactor(s); spk.inmatrix(s); spk.factor(s)
    b = M.rhs_for(A)
    x = b.copy()
    spk.triangularsolve(s, x)
    assert residual(A, x, b) < RESID_TOL
    assert np.allclose(x, np.arange(1, 1601), rtol=1e-9)


def test_csc_interface_refactor_and_pattern_change():
    # test/test_cscinterface.jl:100-201: sparspaklu / sparspaklu! / ldiv! / backslash
    A = M.convdiff3d(7)
    lu = spk.sparspaklu(A)
    b = M.rhs_for(A)
    x = spk.backslash(lu, b)
    assert residual(A, x, b) < RESID_TOL
    A2 = A.copy(); A2.data = A2.data * 1.5                      # same pattern, new values
    order_before = lu.slvr.order.rperm.copy()
    spk.sparspaklu_(lu, A2)
    assert np.array_equal(order_before, lu.slvr.order.rperm)   # ordering + symbolic reused
    x2 = np.zeros_like(b); spk.ldiv(x2, lu, b)
    assert residual(A2, x2, b) < RESID_TOL
    A3 = sp.csc_matrix(A2 + sp.eye(A.shape[0], k=5) * 0.01)     # pattern change
    with pytest.raises(RuntimeError):
        spk.sparspaklu_(lu, A3, allow_pattern_change=False)
    spk.sparspaklu_(lu, A3)
    assert residual(A3, spk.backslash(lu, b), b) < RESID_TOL
    lu0 = spk.sparspaklu(A, factorize=False)
    with pytest.raises(spk.SequenceError):
        spk.triangularsolve(lu0, b.copy())


def test_device_inmatrix_and_refactor():
    """values scattered on the device through the once-built index map; a
    refactorisation uploads nnz(A) values only."""
    A = M.laplacian3d(10)
    s = prepare(A, True, spk.nd_grid_order(10, 10, 10))
    b = s.slvr
    dest, nzval = b._inmatrix_map(A)
    plan = _cudalib.Plan(b)
    plan.inmatrix(nzval, dest)
    assert plan.factor() == 0
    lg = np.zeros(b.lnz.size); plan.get_factors(lg)
    lo, _, _, _ = oracle_factor(b)
    assert rel_err(lg, lo, spd_mask(b)) < FACTOR_RTOL
    plan.inmatrix(nzval * 2.0)                                   # same pattern, map reused
    assert plan.factor() == 0
    plan.get_factors(lg)
    b.lnz *= 2.0
    lo2, _, _, _ = oracle_factor(b)
    assert rel_err(lg, lo2, spd_mask(b)) < FACTOR_RTOL
    plan.destroy()


@pytest.mark.parametrize("spd", [False, True])
def test_multi_rhs(spd):
    """Extension over the reference (single-RHS only, SpkSparseSolver.jl:266): a block of right-hand
    sides; every column must equal the single-RHS solve."""
    A = M.convdiff3d(8) if not spd else M.laplacian3d(8)
    s = prepare(A, spd, spk.nd_grid_order(8, 8, 8))
    b = s.slvr
    plan = _cudalib.Plan(b)
    plan.set_values(b.lnz, None if spd else b.unz); assert plan.factor() == 0
    plan.set_perm(b.order.rperm, b.order.rinvp)
    rng = np.random.default_rng(9876)
    B = np.asfortranarray(rng.random((b.n, 37)))
    X = B.copy(order="F")
    plan.triangularsolve(X)
    for j in (0, 17, 36):
        xj = B[:, j].copy(); plan.triangularsolve(xj)
        assert np.array_equal(xj, X[:, j])
        assert residual(A, X[:, j], B[:, j]) < RESID_TOL
    plan.destroy()


def test_zero_pivot_flag():
    # iflag = -1 on a zero pivot (SpkLUFactor.jl:29-34); the host mirror turns it into an error (SpkSparseBase.jl:386-389)
    A = sp.csc_matrix(np.array([[0.0, 0, 0], [0, 2.0, 1.0], [0, 1.0, 2.0]]))
    A[0, 0] = 0.0
    A = sp.csc_matrix(([0.0, 2.0, 1.0, 1.0, 2.0], ([0, 1, 2, 1, 2], [0, 1, 1, 2, 2])), shape=(3, 3))
    s = prepare(A, False)
    plan = _cudalib.Plan(s.slvr)
    plan.set_values(s.slvr.lnz, s.slvr.unz)
    assert plan.factor() == -1
    with pytest.raises(RuntimeError):
        spk.factor(s)


def test_pivoting_fuzz_on_gpu():
    # test/test_structunsymm.jl:60-90 against the CUDA path: ipiv identical to the oracle's on every case
    rng = np.random.default_rng(4321)
    done = 0
    for k in range(60):
        n = int(rng.integers(4, 40))
        a = sp.random(n, n, density=min(0.9, 3.0 / n + 0.1), random_state=rng, format="csc", data_rvs=rng.random) + sp.identity(n, format="csc")
        a = sp.csc_matrix(a); a.eliminate_zeros()
        if np.linalg.cond(a.toarray()) > 1e8:
            continue
        s = prepare(a, False, maxblocksize=int(rng.integers(2, 12)))
        lo, uo, po, _ = oracle_factor(s.slvr)
        plan, lg, ug, pg, fg = gpu_factor_plan(s.slvr)
        assert np.array_equal(pg, po)
        assert rel_err(lg, lo) < 1e-9 and rel_err(ug, uo) < 1e-9
        bvec = rng.random(n); x = bvec.copy()
        plan.set_perm(s.slvr.order.rperm, s.slvr.order.rinvp); plan.triangularsolve(x)
        assert np.linalg.norm(x - np.linalg.solve(a.toarray(), bvec)) < 1e-9
        plan.destroy(); done += 1
    assert done > 30


def test_dmma_and_dfma_kernels_agree(monkeypatch):
    """The DMMA trailing-update kernels and the small-tile DFMA kernel compute the same factors."""
    A = M.laplacian3d(16)
    s = prepare(A, True, spk.nd_grid_order(16, 16, 16))
    b = s.slvr
    _, l1, _, _, f1 = gpu_factor_plan(b)
    monkeypatch.setenv("SPK_NO_DMMA", "1")
    _, l2, _, _, f2 = gpu_factor_plan(b)
    assert f1 == f2 == 0
    assert rel_err(l1, l2, spd_mask(b)) < 1e-12
    lo, _, _, _ = oracle_factor(b)
    assert rel_err(l1, lo, spd_mask(b)) < FACTOR_RTOL


@pytest.mark.parametrize("spd", [False, True])
def test_large_front_solve_path(spd, monkeypatch):
    """Every front forced onto the per-chunk (multi-block) solve path used for large fronts."""
    monkeypatch.setenv("SPK_SOLVE_SMALL", "0")
    A = M.convdiff3d(10) if not spd else M.laplacian3d(10)
    s = prepare(A, spd, spk.nd_grid_order(10, 10, 10), 6)
    b = s.slvr
    plan, lg, ug, pg, fg = gpu_factor_plan(b)
    bb = M.rhs_for(A)
    x = bb.copy()
    plan.set_perm(b.order.rperm, b.order.rinvp)
    plan.triangularsolve(x)
    assert residual(A, x, bb) < RESID_TOL
    plan.destroy()


VARIANTS = [
    {"SPK_DIAG_SMEM": "1"},                     # shared-memory diagonal-block kernels instead of the register ones
    {"SPK_PANEL_SMEM": "1"},                    # shared-memory panel kernel (both sides for LU)
    {"SPK_DIAG_SMEM": "1", "SPK_PANEL_SMEM": "1", "SPK_PS_WIDTH": "100"},   # panel steps wider than 64 columns
    {"SPK_PIPES": "2"},                         # tree pipelines: two subtree sets on their own stream pairs
    {"SPK_LOOKAHEAD": "0"},                     # one stream
    {"SPK_PDL": "0", "SPK_SOLVE_SMALL": "0"},   # solve steps without programmatic dependent launch
    {"SPK_SOLVE_LNZ": "1"},                     # chunk-by-chunk solve on lnz / unz
    {"SPK_SOLVE_GRAPH": "0", "SPK_SOLVE_SMALL": "0"},
    {"SPK_STORE_OVERLAP": "0", "SPK_LL": "0"},  # one write-back kernel at the end; right-looking in-block updates (LDLt)
    {"SPK_PANEL_REG_MINW": "0", "SPK_PDL_FACTOR": "1"},   # register panel kernel for every width <= 64; PDL in the factorisation
    {"SPK_SOLVE_FLOW": "1", "SPK_FLOW_MIN_STEPS": "1", "SPK_SOLVE_SMALL": "0"},   # dataflow solve sweeps (mailbox-synchronised), every front
    {"SPK_SOLVE_FLOW": "1", "SPK_FLOW_MIN_STEPS": "2", "SPK_PS_WIDTH": "16"},     # dataflow sweeps mixed with the per-step path, many steps
    {"SPK_DMMA_PERSIST": "1", "SPK_GEMM_RESERVE": "16"},  # persistent DMMA blocks with reserved slots
    {"SPK_DMMA_CA": "0", "SPK_DMMA_VARIANT": "6"},        # L2-only operand loads, 4-stage ring
    {"SPK_SOLVE_INV": "0"},                               # in-block triangular solves everywhere (LU default: inverted diagonal blocks)
    {"SPK_SOLVE_INV": "3", "SPK_SOLVE_SMALL": "0"},       # inverted diagonal blocks everywhere (LDL^T default: none)
    {"SPK_SOLVE_INV": "1"}, {"SPK_SOLVE_INV": "2"},
    {"SPK_DMMA_BIG": "1"}, {"SPK_DMMA_VARIANT64": "4"},   # 128 x 64 DMMA tiles / the 16-deep 3-stage pipeline of the 64 x 64 kernel
]


@pytest.mark.parametrize("spd", [False, True])
@pytest.mark.parametrize("env", VARIANTS, ids=["+".join(f"{k[4:]}={v}" for k, v in e.items()) for e in VARIANTS])
def test_kernel_and_schedule_variants_agree(spd, env, monkeypatch):
    """Every alternative kernel / schedule kept behind an environment knob reproduces the default path:
    same pivot sequence, factors within the parity bar, residual within the bar."""
    A = M.convdiff3d(14) if not spd else M.laplacian3d(14)
    s = prepare(A, spd, spk.nd_grid_order(14, 14, 14))
    b = s.slvr
    lo, uo, po, _ = oracle_factor(b)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    plan, lg, ug, pg, fg = gpu_factor_plan(b)
    assert fg == 0
    assert rel_err(lg, lo, spd_mask(b)) < FACTOR_RTOL
    if not spd:
        assert np.array_equal(pg, po)
        assert rel_err(ug, uo) < FACTOR_RTOL
    bb = M.rhs_for(A)
    x = bb.copy()
    plan.set_perm(b.order.rperm, b.order.rinvp)
    plan.triangularsolve(x)
    assert residual(A, x, bb) < RESID_TOL
    plan.destroy()


@pytest.mark.parametrize("spd", [False, True])
def test_float32_twins(spd):
    """spk_*_f32: the Float32 methods of _factor! / _triangularsolve! (the reference sends Float32 to
    sgetrf/sgemm/strsm, SpkSpdMMOps.jl:186-351).  Values cross the boundary as Float32, the arithmetic is the
    FP64 engine's, so the result must match the FP64 oracle run on the same (Float32-representable) input to
    Float32 rounding, with the same pivot sequence."""
    A = (M.convdiff3d(9) if not spd else M.laplacian3d(9)).astype(np.float32).astype(np.float64)
    s = prepare(A, spd, spk.nd_grid_order(9, 9, 9))
    b = s.slvr
    L = _cudalib.lib()
    lo, uo, po, _ = oracle_factor(b)
    lnz = b.lnz.astype(np.float32); unz = b.unz.astype(np.float32); ipiv = np.zeros(b.n, np.int64)
    bb = M.rhs_for(A)
    rhs = np.ascontiguousarray(bb[b.order.rperm - 1]).astype(np.float32)
    eps32 = float(np.finfo(np.float32).eps)
    if spd:
        assert L.spk_ldltfactor_f32(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, lnz) == 0
        assert rel_err(lnz.astype(np.float64), lo, spd_mask(b)) < 4 * eps32
        assert L.spk_ldltsolve_f32(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, rhs) == 1
    else:
        assert L.spk_lufactor_f32(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, ipiv) == 0
        assert np.array_equal(ipiv, po)
        assert rel_err(lnz.astype(np.float64), lo) < 4 * eps32 and rel_err(unz.astype(np.float64), uo) < 4 * eps32
        assert L.spk_lulsolve_f32(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, ipiv, rhs) == 1
        assert L.spk_luusolve_f32(b.n, b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, rhs) == 1
    x = rhs.astype(np.float64)[b.order.rinvp - 1]
    assert residual(A, x, bb) < 200 * eps32         # Float32 factors and right-hand side
    # plan twins
    plan = _cudalib.Plan(b)
    dest, nzval = b._inmatrix_map(A)
    assert L.spk_plan_inmatrix_f32(plan.h, nzval.size, dest.ctypes.data, nzval.astype(np.float32)) == 0
    assert plan.factor() == 0
    plan.set_perm(b.order.rperm, b.order.rinvp)
    x32 = bb.astype(np.float32)
    assert L.spk_plan_triangularsolve_f32(plan.h, x32, 1, b.n) == 0
    assert residual(A, x32.astype(np.float64), bb) < 200 * eps32
    l32 = np.zeros(b.lnz.size, np.float32)
    assert L.spk_plan_get_factors_f32(plan.h, l32.ctypes.data, None, None) == 0
    assert rel_err(l32.astype(np.float64), lo, spd_mask(b) if spd else None) < 4 * eps32
    plan.destroy()


@pytest.mark.parametrize("spd", [False, True])
def test_device_residual_and_refinement(spd):
    """computeresidual (SpkProblem.jl:448-496) and iterative refinement on the device.
    A deliberately poor start (the solution rounded to Float32) must be refined to the FP64 residual bar."""
    A = M.convdiff3d(10) if not spd else M.laplacian3d(10)
    s = prepare(A, spd, spk.nd_grid_order(10, 10, 10))
    b = s.slvr
    plan = _cudalib.Plan(b)
    plan.set_values(b.lnz, None if spd else b.unz); assert plan.factor() == 0
    plan.set_perm(b.order.rperm, b.order.rinvp)
    plan.set_matrix(A)
    rng = np.random.default_rng(2024)
    B = np.asfortranarray(rng.standard_normal((b.n, 3)))
    X = B.copy(order="F"); plan.triangularsolve(X)
    X0 = np.asfortranarray(X.astype(np.float32).astype(np.float64))
    res, rel = plan.residual(B, X0)
    ref = B - A @ X0
    assert np.abs(res - ref).max() <= 1e-13 * np.abs(B).max()    # both sides cancel ~7 digits: compare on the scale of b
    assert np.allclose(rel, np.linalg.norm(ref, axis=0) / np.linalg.norm(B, axis=0), rtol=1e-5)
    assert rel.max() > 1e-9                                     # Float32 rounding of x is visible
    steps, rel2 = plan.refine(B, X0, maxit=4, tol=1e-14)
    assert 1 <= steps <= 4 and rel2.max() <= 1e-14
    for j in range(3):
        assert residual(A, X0[:, j], B[:, j]) < RESID_TOL
    # a converged start needs no correction
    steps, rel3 = plan.refine(B, X0, maxit=4, tol=1e-12)
    assert steps == 0 and rel3.max() <= 1e-12
    plan.destroy()


def test_two_live_plans_of_different_panel_width():
    """Dynamic shared-memory limits are per (kernel, device), not per plan: a plan with narrow
    panel steps created AFTER one with wide steps must not lower the limits under the older plan.  Interleaves
    factor, solve and multi-RHS solve of both."""
    rng = np.random.default_rng(77)
    G = rng.standard_normal((150, 150))
    Awide = sp.csc_matrix(G @ G.T + 150.0 * np.eye(150))         # one dense front, chunks ~75 wide -> wide panel steps
    s1 = prepare(Awide, True, maxblocksize=100)
    A2 = M.convdiff3d(8)
    s2 = prepare(A2, False, spk.nd_grid_order(8, 8, 8), 6)          # narrow panel steps
    p1 = _cudalib.Plan(s1.slvr)
    w1 = p1.stat(10)
    p2 = _cudalib.Plan(s2.slvr)
    assert w1 > 64 and p2.stat(10) < w1
    for plan, s in ((p1, s1), (p2, s2)):
        b = s.slvr
        plan.set_values(b.lnz, None if b.spd else b.unz)
        plan.set_perm(b.order.rperm, b.order.rinvp)
    assert p1.factor() == 0 and p2.factor() == 0
    for rep in range(2):
        for plan, A in ((p1, Awide), (p2, A2), (p1, Awide)):
            B = np.asfortranarray(rng.random((A.shape[0], 9)))
            X = B.copy(order="F"); plan.triangularsolve(X)              # multi-RHS kernels (largest shared memory)
            x = B[:, 0].copy(); plan.triangularsolve(x)
            assert residual(A, x, B[:, 0]) < RESID_TOL
            assert max(residual(A, X[:, j], B[:, j]) for j in range(9)) < RESID_TOL
        p1.set_values(s1.slvr.lnz); assert p1.factor() == 0          # refactor the older plan while the newer one is alive
    lo, _, _, _ = oracle_factor(s1.slvr)
    lg = np.zeros(s1.slvr.lnz.size); p1.get_factors(lg)
    assert rel_err(lg, lo, spd_mask(s1.slvr)) < FACTOR_RTOL
    p1.destroy(); p2.destroy()


def test_solve_without_factors_is_an_error_and_inmatrix_map_guards():
    A = M.convdiff3d(6)                                             # LU: every entry of A has a destination
    s = prepare(A, False, spk.nd_grid_order(6, 6, 6))
    b = s.slvr
    plan = _cudalib.Plan(b)
    plan.set_perm(b.order.rperm, b.order.rinvp)
    with pytest.raises(RuntimeError):
        plan.triangularsolve(M.rhs_for(A))                          # no factors yet
    dest, nzval = b._inmatrix_map(A)
    with pytest.raises(RuntimeError):
        plan.inmatrix(nzval)                                        # no map yet
    plan.inmatrix(nzval, dest)
    with pytest.raises(RuntimeError):
        plan.inmatrix(nzval[:-1])                                   # map was built for another nnz
    # duplicate (i, j) entries accumulate like the reference's `+=`
    dd = np.concatenate([dest, dest[:5]]); vv = np.concatenate([nzval, nzval[:5]])
    plan.inmatrix(vv, dd)
    assert plan.factor() == 0
    Ac = sp.csc_matrix(A); Ac.sort_indices()
    cols = np.repeat(np.arange(Ac.shape[1]), np.diff(Ac.indptr))
    bump = sp.csc_matrix((nzval[:5], (Ac.indices[:5], cols[:5])), shape=A.shape)
    x = M.rhs_for(A); plan.triangularsolve(x)
    assert residual(sp.csc_matrix(A + bump), x, M.rhs_for(A)) < RESID_TOL
    plan.destroy()


def test_condition_estimate_against_dense_cond1():
    """Hager / Higham 1-norm estimate on the resident factors vs numpy.linalg.cond(A, 1).
    The estimator never over-estimates ||inv(A)||_1 and is within a small factor of it (LAPACK's bar: 3)."""
    rng = np.random.default_rng(31)
    mats = [M.laplacian2d(12), M.laplacian3d(7)]
    G = rng.standard_normal((60, 60)); D = np.diag(np.logspace(0, 6, 60))
    mats.append(sp.csc_matrix(G @ D @ G.T + 1e-3 * np.eye(60)))               # ill-conditioned dense SPD
    T = sp.diags([np.full(79, -1.0), np.linspace(2.0, 2.5, 80), np.full(79, -1.0)], [-1, 0, 1], format="csc")
    mats.append(T)
    for A in mats:
        A = sp.csc_matrix(A)
        s = prepare(A, True)
        b = s.slvr
        plan = _cudalib.Plan(b)
        plan.set_values(b.lnz); assert plan.factor() == 0
        plan.set_perm(b.order.rperm, b.order.rinvp); plan.set_matrix(A)
        c, an, ainv, nsolves, lower = plan.condest()
        Ad = A.toarray()
        true_ainv = np.linalg.norm(np.linalg.inv(Ad), 1)
        assert abs(an - np.linalg.norm(Ad, 1)) <= 1e-12 * an
        assert not lower and nsolves <= 11
        assert ainv <= true_ainv * (1 + 1e-8) and ainv >= true_ainv / 3.0
        assert abs(c - an * ainv) <= 1e-12 * c
        plan.destroy()
    # LU plans: a flagged lower bound
    A = M.convdiff3d(6)
    s = prepare(A, False)
    b = s.slvr
    plan = _cudalib.Plan(b)
    plan.set_values(b.lnz, b.unz); assert plan.factor() == 0
    plan.set_perm(b.order.rperm, b.order.rinvp); plan.set_matrix(A)
    c, an, ainv, nsolves, lower = plan.condest()
    true_ainv = np.linalg.norm(np.linalg.inv(A.toarray()), 1)
    assert lower and ainv <= true_ainv * (1 + 1e-8) and ainv >= true_ainv / 10.0
    plan.destroy()


def test_stateless_dropins_reuse_cached_plan_and_resident_factors():
    """The reference's `_triangularsolve!` calls `_lulsolve!` + `_luusolve!` per right-hand side: behind the stateless
    entry points the plan is cached per structure, and a solve that presents the arrays the last factor call wrote
    back uses the resident factors; modified arrays (another address or other values) are uploaded again."""
    import time
    L = _cudalib.lib()
    L.spk_cache_clear()
    A = M.convdiff3d(12)
    s = prepare(A, False, spk.nd_grid_order(12, 12, 12))
    b = s.slvr
    lnz = b.lnz.copy(); unz = b.unz.copy(); ipiv = np.zeros(b.n, np.int64)
    assert L.spk_lufactor_f64(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, ipiv) == 0
    bb = M.rhs_for(A)
    for rep in range(3):                                                  # same arrays: resident factors
        rhs = np.ascontiguousarray((rep + 1.0) * bb[b.order.rperm - 1])
        assert L.spk_lulsolve_f64(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, ipiv, rhs) == 1
        assert L.spk_luusolve_f64(b.n, b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, rhs) == 1
        assert residual(A, rhs[b.order.rinvp - 1], (rep + 1.0) * bb) < RESID_TOL
    # factors of ANOTHER matrix with the same structure, in other arrays: must not be confused with the resident ones
    A2 = A.copy(); A2.data = A2.data * np.linspace(1.0, 2.0, A2.nnz)
    s2 = prepare(A2, False, spk.nd_grid_order(12, 12, 12))
    lo, uo, po, _ = oracle_factor(s2.slvr)
    rhs = np.ascontiguousarray(M.rhs_for(A2)[b.order.rperm - 1])
    assert L.spk_lulsolve_f64(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lo, po, rhs) == 1
    assert L.spk_luusolve_f64(b.n, b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lo, b.xunz, uo, rhs) == 1
    assert residual(A2, rhs[b.order.rinvp - 1], M.rhs_for(A2)) < 1e-11
    # values changed IN PLACE in the arrays of the first factorisation: fingerprint mismatch -> uploaded again
    lnz[:] = lo; unz[:] = uo; ipiv[:] = po
    rhs = np.ascontiguousarray(M.rhs_for(A2)[b.order.rperm - 1])
    assert L.spk_lulsolve_f64(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, ipiv, rhs) == 1
    assert L.spk_luusolve_f64(b.n, b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, rhs) == 1
    assert residual(A2, rhs[b.order.rinvp - 1], M.rhs_for(A2)) < 1e-11
    # SPD twin + a second structure (cache of two plans), then back to the first
    As = M.laplacian3d(9)
    ss = prepare(As, True, spk.nd_grid_order(9, 9, 9))
    ls = ss.slvr.lnz.copy()
    assert L.spk_ldltfactor_f64(ss.slvr.n, ss.slvr.nsuper, ss.slvr.xsuper, ss.slvr.snode, ss.slvr.xlindx, ss.slvr.lindx, ss.slvr.xlnz, ls) == 0
    r2 = np.ascontiguousarray(M.rhs_for(As)[ss.slvr.order.rperm - 1])
    assert L.spk_ldltsolve_f64(ss.slvr.nsuper, ss.slvr.xsuper, ss.slvr.xlindx, ss.slvr.lindx, ss.slvr.xlnz, ls, r2) == 1
    assert residual(As, r2[ss.slvr.order.rinvp - 1], M.rhs_for(As)) < RESID_TOL
    rhs = np.ascontiguousarray(M.rhs_for(A2)[b.order.rperm - 1])
    assert L.spk_lulsolve_f64(b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, ipiv, rhs) == 1
    assert L.spk_luusolve_f64(b.n, b.nsuper, b.xsuper, b.xlindx, b.lindx, b.xlnz, lnz, b.xunz, unz, rhs) == 1
    assert residual(A2, rhs[b.order.rinvp - 1], M.rhs_for(A2)) < 1e-11
    L.spk_cache_clear()
