"""The device schedule (plan.hpp: fronts, relaxed chains, relative indices, panel steps, outer
blocks, solve steps) executed on the host by tests/hostsim must reproduce the oracle."""
import numpy as np
import pytest

import sparspak_jl_b200 as spk
import oracle
from common import CASES, prepare, oracle_factor, spd_mask, rel_err, residual, HostSim, M


@pytest.mark.parametrize("name,build,spd,order,maxblk", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("relax", [(0, 0.0), (4, 0.02)], ids=["fundamental", "relaxed"])
def test_hostsim_matches_oracle(name, build, spd, order, maxblk, relax):
    A = build()
    s = prepare(A, spd, order() if order else None, maxblk)
    b = s.slvr
    lo, uo, po, fo = oracle_factor(b)
    sim = HostSim(b, relax[0], relax[1])
    ls, us, ps, fs = sim.factor()
    assert fs == fo == 0
    nl = int(b.xlnz[b.n]) - 1
    assert rel_err(ls, lo[:nl], spd_mask(b)[:nl]) < 1e-12
    if not spd:
        assert np.array_equal(ps, po)
        assert rel_err(us, uo) < 1e-12
    bb = M.rhs_for(A)
    rhs = np.ascontiguousarray(bb[b.order.rperm - 1])
    x = sim.solve(rhs)[b.order.rinvp - 1]
    assert residual(A, x, bb) < 1e-13


def test_hostsim_many_children_tail_path():
    # arrow matrix: the root front has n-1 children -> exercises the assemble tail kernel path
    import scipy.sparse as sp
    n = 40
    A = sp.lil_matrix((n, n)); A.setdiag(4.0); A[n - 1, :] = -0.1; A[:, n - 1] = -0.1; A[n - 1, n - 1] = 10.0
    A = sp.csc_matrix(A)
    for spd in (False, True):
        s = prepare(A, spd)
        lo, uo, po, _ = oracle_factor(s.slvr)
        ls, us, ps, _ = HostSim(s.slvr).factor()
        nl = int(s.slvr.xlnz[n]) - 1
        assert rel_err(ls, lo[:nl], spd_mask(s.slvr)[:nl]) < 1e-13


def test_plan_statistics_cfg1():
    # config 1 (2-D 5-point 100x100, SPD, nested dissection): structure figures used by DESIGN.md
    A = M.laplacian2d(100)
    s = spk.SparseSpdSolver(A)
    spk.findorder(s, spk.nd_grid_order(100, 100)); spk.symbolicfactor(s)
    sim = HostSim(s.slvr, alloc=False)
    assert s.slvr.n == 10000
    assert sim.stat(0) < s.slvr.nsuper            # relaxed chains merge supernodes into fewer fronts
    assert sim.stat(1) < 40                       # few front levels


@pytest.mark.parametrize("spd", [False, True])
def test_hostsim_large_front_solve_path(spd, monkeypatch):
    # force every front onto the per-chunk (multi-block) solve path used for large fronts
    monkeypatch.setenv("SPK_SOLVE_SMALL", "0")
    A = M.convdiff3d(8) if not spd else M.laplacian3d(8)
    s = prepare(A, spd, spk.nd_grid_order(8, 8, 8), 6)
    b = s.slvr
    sim = HostSim(b)
    sim.factor()
    bb = M.rhs_for(A)
    rhs = np.ascontiguousarray(bb[b.order.rperm - 1])
    x = sim.solve(rhs)[b.order.rinvp - 1]
    assert residual(A, x, bb) < 1e-13


SCHED_KNOBS = [
    {"SPK_OB_STEPS": "2", "SPK_PS_WIDTH": "8"},                                   # many outer blocks: strip / rest / delayed updates
    {"SPK_OB_STEPS": "2", "SPK_PS_WIDTH": "8", "SPK_SPLIT_REST": "1"},            # delayed update in two launches (early part first)
    {"SPK_OB_STEPS": "1", "SPK_PS_WIDTH": "16", "SPK_SPLIT_REST": "1", "SPK_DMMA_NARROW": "0"},
    {"SPK_OB_STEPS": "3", "SPK_PS_WIDTH": "8", "SPK_DMMA_BIG": "1", "SPK_DMMA_NARROW": "100000"},   # 128-row and 64 x 32 tile lists
]


@pytest.mark.parametrize("spd", [False, True])
@pytest.mark.parametrize("env", SCHED_KNOBS, ids=["+".join(f"{k[4:]}={v}" for k, v in e.items()) for e in SCHED_KNOBS])
def test_hostsim_schedule_knobs(spd, env, monkeypatch):
    """The launch lists built under the schedule / tile-shape knobs (split delayed update, 128 x 64 / 64 x 64 / 64 x 32
    DMMA tile lists) are complete and correctly ordered: executed serially on the host they reproduce the oracle."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = 10
    A = M.convdiff3d(g) if not spd else M.laplacian3d(g)
    s = prepare(A, spd, spk.nd_grid_order(g, g, g))
    b = s.slvr
    lo, uo, po, fo = oracle_factor(b)
    ls, us, ps, fs = HostSim(b).factor()
    assert fs == fo == 0
    nl = int(b.xlnz[b.n]) - 1
    assert rel_err(ls, lo[:nl], spd_mask(b)[:nl]) < 1e-12
    if not spd:
        assert np.array_equal(ps, po)
        assert rel_err(us, uo) < 1e-12


def test_position_maps_do_not_depend_on_host_threads(monkeypatch):
    """analyze() builds the per-chunk position maps with host threads over disjoint front ranges once they are
    large (> 4 Mi entries): the result must be the serial one for every thread count."""
    g = 48
    A = M.laplacian3d(g)
    s = prepare(A, True, spk.nd_grid_order(g, g, g))
    fp = set()
    for nth in ("1", "2", "7", "33"):
        monkeypatch.setenv("SPK_HOST_THREADS", nth)
        sim = HostSim(s.slvr, alloc=False)
        assert sim.stat(21) > (1 << 22)              # large enough for the threaded path
        fp.add(sim.stat(20))
    assert len(fp) == 1
