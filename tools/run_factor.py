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
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
g = a.grid
A = spk.matrices.laplacian3d(g) if a.kind == "spd" else spk.matrices.convdiff3d(g)
s = (spk.SparseSpdSolver if a.kind == "spd" else spk.SparseSolver)(A)
spk.findorder(s, spk.nd_grid_order(g, g, g)); spk.symbolicfactor(s)
b = s.slvr
dest, nzval = b._inmatrix_map(A)
plan = _cudalib.Plan(b)
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
        kinds = ["asm", "asm_tail", "diag", "panel", "gemm_small", "gemm_dmma64", "gemm_dmma128"]
        print("   " + "  ".join(f"{k}={plan.statf(10 + i):.2f}ms/{int(plan.statf(30 + i))}" for i, k in enumerate(kinds)))
        print(f"   dmma: {plan.statf(4) / max(plan.statf(5), 1e-9) / 1e9:.2f} TFLOP/s")
bb = spk.matrices.rhs_for(A)
for r in range(a.solve):
    x = bb.copy(); plan.triangularsolve(x)
    print(f"solve {r}: {plan.statf(3):.2f} ms, launches {plan.stat(1)}, residual {np.linalg.norm(A @ x - bb) / np.linalg.norm(bb):.2e}", flush=True)
