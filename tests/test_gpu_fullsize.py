"""BASELINE.json-size checks through size-independent properties (the oracle would take too long):
solve residual, linearity of the solve, refactor idempotence, and pivot identity on the
diagonally dominant LU config."""
import numpy as np
import pytest

import sparspak_jl_b200 as spk
from sparspak_jl_b200 import _cudalib
from common import M, residual, RESID_TOL

pytestmark = pytest.mark.gpu


def _factor(A, spd, order):
    s = (spk.SparseSpdSolver if spd else spk.SparseSolver)(A)
    spk.findorder(s, order); spk.symbolicfactor(s)
    b = s.slvr
    dest, nzval = b._inmatrix_map(A)
    plan = _cudalib.Plan(b)
    plan.inmatrix(nzval, dest)
    assert plan.factor() == 0
    plan.set_perm(b.order.rperm, b.order.rinvp)
    return s, plan, dest, nzval


def test_cfg2_laplacian_64_cubed_spd():
    # config 2: 3-D 7-point Laplacian 64^3 (n = 262,144), LDL^T, nested dissection
    g = 64
    A = M.laplacian3d(g)
    s, plan, dest, nzval = _factor(A, True, spk.nd_grid_order(g, g, g))
    b = M.rhs_for(A)
    x = b.copy(); plan.triangularsolve(x)
    assert residual(A, x, b) < RESID_TOL
    # linearity: solve(2b + c) == 2 solve(b) + solve(c)
    rng = np.random.default_rng(9876)
    c = rng.random(A.shape[0]); xc = c.copy(); plan.triangularsolve(xc)
    y = 2 * b + c; plan.triangularsolve(y)
    assert np.linalg.norm(y - (2 * x + xc)) / np.linalg.norm(y) < 1e-11
    # refactor with the same pattern is idempotent (bit-identical factors)
    l1 = np.zeros(s.slvr.lnz.size); plan.get_factors(l1)
    plan.inmatrix(nzval); assert plan.factor() == 0
    l2 = np.zeros(s.slvr.lnz.size); plan.get_factors(l2)
    assert np.array_equal(l1, l2)
    plan.destroy()


def test_cfg3_like_convdiff_48_cubed_lu():
    # config 3 at a reduced grid (48^3): upwind convection-diffusion, LU; diagonally dominant => ipiv[k] = k
    g = 48
    A = M.convdiff3d(g)
    s, plan, dest, nzval = _factor(A, False, spk.nd_grid_order(g, g, g))
    b = M.rhs_for(A)
    x = b.copy(); plan.triangularsolve(x)
    assert residual(A, x, b) < RESID_TOL
    ipiv = np.zeros(s.slvr.n, np.int64)
    plan.get_factors(None, None, ipiv)
    sb = s.slvr
    local = np.arange(sb.n) - (sb.xsuper[sb.snode - 1] - 1) + 1
    assert np.array_equal(ipiv, local)
    plan.destroy()


def test_cfg1_laplacian2d_100():
    # config 1: 2-D 5-point Laplacian 100x100, LDL^T, nested dissection (the reference's CPU-runnable case)
    A = M.laplacian2d(100)
    s, plan, dest, nzval = _factor(A, True, spk.nd_grid_order(100, 100))
    b = M.rhs_for(A)
    x = b.copy(); plan.triangularsolve(x)
    assert residual(A, x, b) < RESID_TOL
    plan.destroy()
