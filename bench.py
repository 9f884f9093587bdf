#!/usr/bin/env python
"""bench.py — BASELINE.json metric: FP64 factor GFLOP/s (+ factor / solve seconds) of the
supernodal LDL^T of a 3-D 7-point Laplacian (96^3 by default), nested-dissection order.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --steps K --warmup W    (reference arm: CPU oracle + OpenBLAS)

A "step" = one numeric factorisation (values re-scattered into the fronts, factor, factors
written back in the reference layout) + one triangular solve.  `value` is measured with the
matrix values / rhs already resident in HBM; `e2e` goes through the plan C-ABI with HOST
buffers (H2D of nnz(A) values and the rhs, D2H of the solution, inside the timed region).
Ordering / symbolic factorisation are host code in the reference too and are not timed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- problem
def build_problem(grid, kind="spd"):
    import sparspak_jl_b200 as spk
    t0 = time.time()
    if kind == "spd":
        A = spk.matrices.laplacian3d(grid); s = spk.SparseSpdSolver(A)
    else:
        A = spk.matrices.convdiff3d(grid); s = spk.SparseSolver(A)
    spk.findorder(s, spk.nd_grid_order(grid, grid, grid))
    spk.symbolicfactor(s)
    dest, nzval = s.slvr._inmatrix_map(A)
    log(f"[bench] {kind} {grid}^3: n={s.slvr.n} nsuper={s.slvr.nsuper} nnz(lnz)={int(s.slvr.xlnz[-1]) - 1:.3e} "
        f"host analysis {time.time() - t0:.1f}s")
    return spk, A, s, dest, nzval


def structural_flops(b):
    cc = (np.repeat(np.diff(b.xlindx), np.diff(b.xsuper)) - (np.arange(b.n) - (b.xsuper[b.snode - 1] - 1))).astype(np.float64)
    s1, s2 = cc.sum(), (cc * cc).sum()
    return (s2 if b.spd else 2 * s2 - s1), s1


# --------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_reference_run(grid, steps, warmup):
    """The reference's CPU path = the oracle restatement with its dense call sites on OpenBLAS
    (what Sparspak.jl executes for Float64, SpkSpdMMOps.jl:222-351), all host threads."""
    import oracle
    spk, A, s, dest, nzval = build_problem(grid)
    b = s.slvr
    spk.inmatrix(s)
    oracle.use_openblas(True)
    F, nnzl = structural_flops(b)
    tf, ts = [], []
    rhs0 = np.ascontiguousarray(spk.matrices.rhs_for(A)[b.order.rperm - 1])
    for it in range(warmup + steps):
        lnz = b.lnz.copy()
        t0 = time.perf_counter(); fl = oracle.ldltfactor(b, lnz); t1 = time.perf_counter()
        rhs = rhs0.copy()
        oracle.ldltsolve(b, lnz, rhs); t2 = time.perf_counter()
        assert fl == 0
        if it >= warmup:
            tf.append(t1 - t0); ts.append(t2 - t1)
    x = rhs[b.order.rinvp - 1]
    bb = spk.matrices.rhs_for(A)
    res = float(np.linalg.norm(A @ x - bb) / np.linalg.norm(bb))
    return dict(flops=F, factor_s=float(np.mean(tf)), solve_s=float(np.mean(ts)), residual=res, n=b.n)


def measure_dgemm_peak(torch, n=8192, reps=5):
    """FP64 tensor peak stand-in: cuBLAS DGEMM n^3 (MEASURED_PEAKS.json has no FP64 entry)."""
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def main():
    # Libraries (NCCL prints its version banner) must not pollute stdout: the contract is ONE JSON line there.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    try:
        _main(json_out)
    finally:
        json_out.flush()


def _main(json_out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--grid", type=int, default=int(os.environ.get("SPK_BENCH_GRID", "96")))
    ap.add_argument("--cpu-grid", type=int, default=int(os.environ.get("SPK_BENCH_CPU_GRID", "48")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel-kind time breakdown to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    metric = "fp64_factor_gflops"
    workload = f"3D 7-point Laplacian {args.grid}^3 SPD supernodal LDL^T, nested dissection, factor + 1-RHS solve"

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count()
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
        r = cpu_reference_run(args.cpu_grid, max(args.steps, 1), min(args.warmup, 1))
        gf = r["flops"] / r["factor_s"] / 1e9
        sample = f"3D 7-point Laplacian {args.cpu_grid}^3 (bounded sample of the {args.grid}^3 workload), oracle restatement + OpenBLAS"
        out = {"impl": "reference", "metric": metric, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": (r["factor_s"] + r["solve_s"]) * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload, "sample": sample, "grid": args.cpu_grid},
               "factor_s": r["factor_s"], "solve_s": r["solve_s"], "residual": r["residual"],
               "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": "port", "sample": sample},
               "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out), file=json_out)
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from sparspak_jl_b200 import _cudalib

    spk, A, s, dest, nzval = build_problem(args.grid)
    b = s.slvr
    F, nnzl = structural_flops(b)
    bb = spk.matrices.rhs_for(A)
    rhs_perm = torch.from_numpy(np.ascontiguousarray(bb[b.order.rperm - 1])).cuda()
    work = rhs_perm.clone()
    if world == 1:
        plan = _cudalib.Plan(b, device=local_rank)
        ds = None
    else:
        from sparspak_jl_b200.multigpu import CudaEngine, DistributedSolver
        eng = CudaEngine(b, rank, world, local_rank)
        plan = eng.plan
        ds = DistributedSolver(eng, rank, world)
    plan.set_perm(b.order.rperm, b.order.rinvp)
    plan.inmatrix(nzval, dest)                      # values resident from here on
    log(f"[bench] rank {rank}: fronts={plan.stat(2)} levels={plan.stat(3)} arena={plan.stat(6) * 8 / 2**30:.1f} GiB "
        f"factor launches={plan.stat(7)} solve launches={plan.stat(8)} top-set fronts={plan.stat(13)} structural flops={F:.3e}")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step_resident():
        """one numeric factorisation + one solve, inputs resident in HBM; returns (flag, factor ms, solve ms)"""
        if ds is None:
            plan.reassemble()
            fl = plan.factor()
            work.copy_(rhs_perm); torch.cuda.current_stream().synchronize()
            plan.solve_device(work.data_ptr(), 1, b.n, 0)
            return fl, plan.statf(2), plan.statf(3)
        sync_all(); t0 = time.perf_counter()
        plan.reassemble()
        fl = ds.factor()                             # phase 0, broadcast of subtree-root fronts, phase 1
        torch.cuda.synchronize(); t1 = time.perf_counter()
        work.copy_(rhs_perm)
        ds.solve(work)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        return fl, (t1 - t0) * 1e3, (t2 - t1) * 1e3

    for _ in range(warmup):
        fl, _, _ = step_resident()
        assert fl == 0
    sampler = ClockSampler(local_rank); sampler.start()
    sync_all()
    t0 = time.perf_counter()
    f_ms, s_ms = [], []
    for _ in range(args.steps):
        fl, fm, sm = step_resident()
        f_ms.append(fm); s_ms.append(sm)
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    factor_ms, solve_ms = float(np.mean(f_ms)), float(np.mean(s_ms))
    launches = (plan.stat(0) + plan.stat(1) + 2) * args.steps
    # correctness of what was timed
    x = work.cpu().numpy()[b.order.rinvp - 1]
    resid = float(np.linalg.norm(A @ x - bb) / np.linalg.norm(bb))

    # e2e: plan C-ABI with HOST buffers (H2D of nnz(A) values + rhs, D2H of the solution, in the timed region)
    e_ms = []
    # host buffers in pinned memory (what a caller that cares about transfer time registers once)
    nz_host = torch.from_numpy(np.ascontiguousarray(nzval)).pin_memory()
    xb_host = torch.from_numpy(bb.copy()).pin_memory()
    nzval_h, xb = nz_host.numpy(), xb_host.numpy()
    pinned = torch.from_numpy(np.ascontiguousarray(bb[b.order.rperm - 1])).pin_memory()
    for it in range(1 + args.steps):
        sync_all(); t1 = time.perf_counter()
        plan.inmatrix(nzval_h)                       # H2D nnz(A) doubles
        if ds is None:
            fl = plan.factor()
            xb[:] = bb
            plan.triangularsolve(xb)                 # H2D + D2H n doubles
        else:
            fl = ds.factor()
            work.copy_(pinned, non_blocking=False)   # H2D n doubles
            ds.solve(work)
            xb[:] = work.cpu().numpy()[b.order.rinvp - 1]   # D2H n doubles
        sync_all(); t2 = time.perf_counter()
        if it > 0:
            e_ms.append((t2 - t1) * 1e3)
    e2e_ms = float(np.mean(e_ms))
    e2e_resid = float(np.linalg.norm(A @ xb - bb) / np.linalg.norm(bb))

    # roofline of the dominant kernel (DMMA trailing update): per-launch CUDA events on the plan's stream
    plan.stat(100)
    plan.reassemble()
    if ds is None:
        plan.factor()
    else:
        ds.factor()
    gemm_flops, gemm_ms = plan.statf(4), plan.statf(5)
    kinds = ["asm", "asm_tail", "diag", "panel", "gemm_small", "gemm_dmma64", "gemm_dmma128"]
    breakdown = {k: {"ms": plan.statf(10 + i), "launches": int(plan.statf(30 + i))} for i, k in enumerate(kinds)}
    prof_total = plan.statf(2)
    plan.stat(101)
    if args.profile:
        log("[bench] profiled factor (per-launch events, one stream): total %.2f ms" % prof_total)
        for k, v in breakdown.items():
            log(f"    {k:14s} {v['ms']:9.3f} ms  {v['launches']:6d} launches")
    phase_ms = (plan.statf(6), plan.statf(7)) if ds is not None else None
    plan.destroy()
    dgemm_tf = measure_dgemm_peak(torch)
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    tmax = torch.tensor([factor_ms, solve_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    factor_ms, solve_ms, e2e_ms = [float(v) for v in tmax.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_rep = 1                                        # N>1: ONE factorisation partitioned by elimination subtrees (strong scaling)
    value = n_rep * F / (factor_ms * 1e-3) / 1e9
    out = {
        "metric": metric, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "grid": args.grid, "n": int(b.n), "nnz_lnz": int(b.xlnz[-1]) - 1,
                   "structural_flops": F, "ordering": "geometric nested dissection (harness callback)",
                   "l2": "working set (frontal arena + factors, tens of GB) far exceeds the 126 MB L2; no flush needed",
                   "step": "arena clear + device scatter of A's values + numeric factorisation + one triangular solve; "
                           "value = structural factor flops / factor_s (BASELINE metric: factor GFLOP/s), "
                           "ms_per_step = the whole step, solve_s reported separately",
                   "parallelism": (f"{world} elimination subtrees -> GPUs, NCCL broadcast of subtree-root fronts, replicated top set"
                                   if world > 1 else "single GPU")},
        "factor_s": factor_ms * 1e-3, "solve_s": solve_ms * 1e-3, "residual": resid,
        "e2e": {"value": n_rep * F / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s (structural factor flops / wall time of inmatrix+factor+solve through the plan C-ABI)",
                "ms": e2e_ms, "h2d_bytes_per_step": int(nzval.nbytes + bb.nbytes), "d2h_bytes_per_step": int(bb.nbytes),
                "residual": e2e_resid},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": dgemm_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / dgemm_tf if dgemm_tf > 0 else None, "traffic": None,
                     "kernel": "k_gemm_dmma (DMMA trailing update)", "peak_source": "cuBLAS DGEMM 8192^3 measured in this run",
                     "kernel_flops": gemm_flops, "kernel_ms": gemm_ms, "share_of_factor": gemm_ms / prof_total if prof_total else None},
        "breakdown_ms": breakdown, "phase_ms": phase_ms,
        "wall_s_timed_region": wall,
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count()
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
        r = cpu_reference_run(args.cpu_grid, 1, 0)
        out["cpu_baseline"] = {"value": r["flops"] / r["factor_s"] / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                               "sample": f"3D 7-point Laplacian {args.cpu_grid}^3 LDL^T factor, oracle restatement + OpenBLAS "
                                         f"({r['factor_s']:.2f} s factor, {r['solve_s']:.3f} s solve)"}
    print(json.dumps(out), file=json_out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
