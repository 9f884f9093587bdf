"""Profiling target: build the plan for a grid, run `--reps` numeric factorisations (+ solves).
Used under ncu (see profiles/README.md); never a bench number."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sparspak_jl_b200 as spk
from sparspak_jl_b200 import _cudalib

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=96)
ap.add_argument("--kind", default="spd")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--solve", type=int, default=0)
ap.add_argument("--nrhs", type=int, default=1)
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
g = a.grid
import time
t0 = time.time()
dof = 1
if a.kind == "spd":
    A = spk.matrices.laplacian3d(g)
elif a.kind == "lu":
    A = spk.matrices.convdiff3d(g)
else:                                   # "el": config 5, 27-point stencil x 3 dof (LDL^T)
    A = spk.matrices.elasticity27(g); dof = 3
s = (spk.SparseSolver if a.kind == "lu" else spk.SparseSpdSolver)(A)
spk.findorder(s, spk.nd_grid_order(g, g, g, dof)); spk.symbolicfactor(s)
print(f"host analysis {time.time() - t0:.1f}s: n={s.slvr.n} nnz(A)={A.nnz} nsuper={s.slvr.nsuper} nnz(lnz)={int(s.slvr.xlnz[-1]) - 1:.3e}", flush=True)
b = s.slvr
dest, nzval = b._inmatrix_map(A)
plan = _cudalib.Plan(b)
print(f"plan: fronts={plan.stat(2)} levels={plan.stat(3)} arena={plan.stat(6) * 8 / 2**30:.1f} GiB flops={plan.statf(0):.3e}", flush=True)
plan.set_perm(b.order.rperm, b.order.rinvp)
plan.inmatrix(nzval, dest)
if a.profile:
    plan.stat(100)
for r in range(a.reps):
    if r:
        plan.reassemble()
    fl = plan.factor()
    print(f"factor {r}: flag {fl} {plan.statf(2):.2f} ms, {plan.statf(0) / plan.statf(2) / 1e6:.1f} GFLOP/s, launches {plan.stat(0)}", flush=True)
    if a.profile:
        kinds = ["asm", "asm_tail", "diag", "panel", "gemm_small", "gemm_dmma_128x64", "gemm_dmma_64x64"]
        print("   " + "  ".join(f"{k}={plan.statf(10 + i):.2f}ms/{int(plan.statf(30 + i))}" for i, k in enumerate(kinds)))
        print(f"   dmma: {plan.statf(4) / max(plan.statf(5), 1e-9) / 1e9:.2f} TFLOP/s")
bb = spk.matrices.rhs_for(A)
for r in range(a.solve):
    if a.nrhs == 1:
        x = bb.copy(); plan.triangularsolve(x)
        res = np.linalg.norm(A @ x - bb) / np.linalg.norm(bb)
    else:
        rng = np.random.default_rng(9876)
        B = np.asfortranarray(rng.random((b.n, a.nrhs)))
        X = B.copy(order="F"); plan.triangularsolve(X)
        res = max(np.linalg.norm(A @ X[:, j] - B[:, j]) / np.linalg.norm(B[:, j]) for j in (0, a.nrhs // 2, a.nrhs - 1))
    print(f"solve {r}: {plan.statf(3):.2f} ms ({a.nrhs} rhs), launches {plan.stat(1)}, residual {res:.2e}", flush=True)
