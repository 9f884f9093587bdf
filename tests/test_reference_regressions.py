"""The reference's FIXED regression inputs (tests/golden/reference_matrices.json, transcribed by
tests/golden/make_reference_matrices.py from test/test_structunsymm.jl:13-60, test/test_small.jl:35-149,
test/test_sparse_method.jl:389-424) through the oracle and the host simulator (CPU), and through the CUDA
path (-m gpu).  The reference's own bar is ||x - A\\b|| / ||A\\b|| < 1e-6; here the solves must also meet the
1e-12 residual bar and the CUDA factors / pivots must equal the oracle's."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import sparspak_jl_b200 as spk
import oracle
from common import prepare, oracle_factor, rel_err, residual, HostSim, FACTOR_RTOL, RESID_TOL

FIX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_matrices.json")))
NAMES = [k for k in FIX if not k.startswith("_")]


def _matrix(name):
    f = FIX[name]
    A = sp.csc_matrix((f["V"], (np.array(f["I"]) - 1, np.array(f["J"]) - 1)), shape=(f["n"], f["n"]))
    return A, np.array(f["rhs"], dtype=np.float64)


def _problem(name):
    """entered the way the reference tests do: element by element (inaij!) + rhs (inbi!)"""
    f = FIX[name]
    p = spk.Problem(f["n"], f["n"])
    for i, j, v in zip(f["I"], f["J"], f["V"]):
        spk.inaij(p, i, j, v)
    for i, v in enumerate(f["rhs"]):
        spk.inbi(p, i + 1, v)
    return p


@pytest.mark.parametrize("name", NAMES)
def test_oracle_and_hostsim_on_reference_regressions(name):
    A, rhs = _matrix(name)
    s = prepare(_problem(name).csc(), False)            # MMD, as in the reference tests
    b = s.slvr
    lo, uo, po, fo = oracle_factor(b)
    assert fo == 0
    x = oracle.triangularsolve(b, lo, uo, po, rhs)
    xd = np.linalg.solve(A.toarray(), rhs)
    assert np.linalg.norm(x - xd) / np.linalg.norm(xd) < 1e-12       # reference bar: 1e-6
    ls, us, ps, fs = HostSim(b).factor()
    assert fs == 0 and np.array_equal(ps, po)
    assert rel_err(ls, lo) < 1e-13 and rel_err(us, uo) < 1e-13


def test_31x31_is_the_matrix_the_reference_holds():
    A, rhs = _matrix("sparse_method_31x31")
    assert A.shape == (31, 31) and A.nnz == 95
    assert np.all(A.diagonal() == 20.0) and abs(A - A.T).max() == 0.0
    assert rhs.tolist() == list(range(1, 32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_on_reference_regressions(name):
    from sparspak_jl_b200 import _cudalib
    A, rhs = _matrix(name)
    p = _problem(name)
    s = prepare(p.csc(), False)
    b = s.slvr
    lo, uo, po, _ = oracle_factor(b)
    plan = _cudalib.Plan(b)
    plan.set_values(b.lnz, b.unz)
    assert plan.factor() == 0
    lg = np.zeros(b.lnz.size); ug = np.zeros(b.unz.size); pg = np.zeros(b.n, np.int64)
    plan.get_factors(lg, ug, pg)
    assert np.array_equal(pg, po)
    assert rel_err(lg, lo) < FACTOR_RTOL and rel_err(ug, uo) < FACTOR_RTOL
    x = rhs.copy()
    plan.set_perm(b.order.rperm, b.order.rinvp); plan.triangularsolve(x)
    xd = np.linalg.solve(A.toarray(), rhs)
    assert np.linalg.norm(x - xd) / np.linalg.norm(xd) < 1e-12
    assert residual(A, x, rhs) < RESID_TOL
    plan.destroy()
    # and the reference's own call sequence: solve!(SparseSolver(p))
    s2 = spk.SparseSolver(p)
    assert spk.solve(s2)
    assert np.linalg.norm(p.x - xd) / np.linalg.norm(xd) < 1e-12
