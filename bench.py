#!/usr/bin/env python
"""bench.py — BASELINE.json metric: FP64 factor GFLOP/s (+ factor / solve seconds) of the supernodal
factorisation of a synthetic grid matrix (default: config 4, LDL^T of the 96^3 7-point Laplacian,
nested-dissection order) on 1/2/4/8 B200, beside the reference's CPU path on the host cores.

  python bench.py --gpus N --steps K --warmup W [--config cfg1..cfg5]      (our arm)
  python bench.py --impl reference --steps K --warmup W [--config ...]     (reference arm: the SAME matrix through
                                                                            the reference schedule + OpenBLAS on the host)

A "step" = one numeric refactorisation (frontal arena cleared, A's values re-scattered on the device, factor,
factors written back in the reference layout) + one triangular solve (config 5: 128 right-hand sides).
`value` = structural factor flops / factor seconds (clear + scatter + factor + write-back, CUDA events on the
plan's stream; N > 1: wall clock between barriers, max over ranks) with A's values / rhs resident in HBM;
`e2e` = the same through the plan C-ABI with HOST buffers (H2D of nnz(A) values and the rhs, D2H of the
solution inside the timed region).  Ordering / symbolic factorisation are host code in the reference too and
are not timed.  Both arms print the same `config`, `metric`, `unit` and `e2e.unit`.
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

METRIC = "fp64_factor_gflops"
UNIT = "GFLOP/s"

# BASELINE.json `configs`, in order.  `sample_grid`: the bounded CPU sample our arm's `cpu_baseline` leg runs.
CONFIGS = {
    "cfg1": dict(kind="spd", matrix="lap2d", grid=100, dof=1, nrhs=1, sample_grid=100,
                 workload="2D 5-point Laplacian {g}x{g} SPD supernodal LDL^T, nested dissection, factor + 1-RHS solve"),
    "cfg2": dict(kind="spd", matrix="lap3d", grid=64, dof=1, nrhs=1, sample_grid=48,
                 workload="3D 7-point Laplacian {g}^3 SPD supernodal LDL^T, nested dissection, factor + 1-RHS solve"),
    "cfg3": dict(kind="lu", matrix="convdiff", grid=80, dof=1, nrhs=1, sample_grid=40,
                 workload="3D upwind convection-diffusion {g}^3 supernodal LU with in-supernode partial pivoting, nested dissection, factor + 1-RHS solve"),
    "cfg4": dict(kind="spd", matrix="lap3d", grid=96, dof=1, nrhs=1, sample_grid=64,
                 workload="3D 7-point Laplacian {g}^3 SPD supernodal LDL^T, nested dissection, factor + 1-RHS solve"),
    "cfg5": dict(kind="spd", matrix="elasticity", grid=64, dof=3, nrhs=128, sample_grid=24,
                 workload="3D 27-point 3-dof elasticity-like {g}^3 SPD supernodal LDL^T, nested dissection, refactor (same pattern) + 128-RHS solve"),
}


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
def build_problem(cfg, grid):
    import sparspak_jl_b200 as spk
    t0 = time.time()
    M = spk.matrices
    if cfg["matrix"] == "lap2d":
        A = M.laplacian2d(grid); order = spk.nd_grid_order(grid, grid)
    elif cfg["matrix"] == "lap3d":
        A = M.laplacian3d(grid); order = spk.nd_grid_order(grid, grid, grid)
    elif cfg["matrix"] == "convdiff":
        A = M.convdiff3d(grid); order = spk.nd_grid_order(grid, grid, grid)
    else:
        A = M.elasticity27(grid); order = spk.nd_grid_order(grid, grid, grid, cfg["dof"])
    s = (spk.SparseSpdSolver if cfg["kind"] == "spd" else spk.SparseSolver)(A)
    spk.findorder(s, order)
    spk.symbolicfactor(s)
    dest, nzval = s.slvr._inmatrix_map(A)
    log(f"[bench] {cfg['matrix']} {cfg['kind']} grid {grid}: n={s.slvr.n} nsuper={s.slvr.nsuper} "
        f"nnz(lnz)={int(s.slvr.xlnz[-1]) - 1:.3e} host analysis {time.time() - t0:.1f}s")
    return spk, A, s, dest, nzval


def structural_flops(b):
    cc = (np.repeat(np.diff(b.xlindx), np.diff(b.xsuper)) - (np.arange(b.n) - (b.xsuper[b.snode - 1] - 1))).astype(np.float64)
    s1, s2 = cc.sum(), (cc * cc).sum()
    return (s2 if b.spd else 2 * s2 - s1), s1


def make_rhs(spk, A, nrhs):
    """1 RHS: b = A * (1..n) (makerhs!, SpkProblem.jl:408-412); a block: default_rng(9876)."""
    if nrhs == 1:
        return spk.matrices.rhs_for(A)
    return np.asfortranarray(np.random.default_rng(9876).random((A.shape[0], nrhs)))


def rel_residual(A, x, b):
    r = A @ x - b
    return float(np.max(np.linalg.norm(r.reshape(b.shape[0], -1), axis=0) / np.linalg.norm(b.reshape(b.shape[0], -1), axis=0)))


def describe_config(name, cfg, grid, b, F):
    """The workload description: IDENTICAL in both arms (the driver compares it)."""
    return {"workload": cfg["workload"].format(g=grid), "name": name, "grid": int(grid), "n": int(b.n),
            "nnz_lnz": int(b.xlnz[-1]) - 1, "structural_flops": float(F), "nrhs": int(cfg["nrhs"]),
            "ordering": "geometric nested dissection (harness callback through findorder!(s, orderfunction))",
            "l2": "working set (frontal arena + factors, GBs) far exceeds the 126 MB L2; no flush needed",
            "step": "numeric refactorisation (A's values -> factor storage, factor, factors in the reference layout) + "
                    "triangular solve; value = structural factor flops / factor seconds"}


# --------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_reference_run(cfg, grid, steps, warmup, budget_s=None):
    """The reference's CPU path = the oracle restatement with its dense call sites on OpenBLAS
    (what Sparspak.jl executes for Float64, SpkSpdMMOps.jl:222-351), all host threads.
    Runs `warmup` untimed + up to `steps` timed repetitions; stops early once `budget_s` is spent (>= 1 timed)."""
    import oracle
    spk, A, s, dest, nzval = build_problem(cfg, grid)
    b = s.slvr
    spk.inmatrix(s)
    oracle.use_openblas(True)
    F, nnzl = structural_flops(b)
    spd = b.spd
    tf, ts = [], []
    bb = make_rhs(spk, A, cfg["nrhs"])
    bcols = bb.reshape(b.n, -1, order="F")
    t_start = time.perf_counter()
    it = 0
    while True:
        lnz = b.lnz.copy(); unz = b.unz.copy(); ipiv = np.zeros(b.n, np.int64)
        t0 = time.perf_counter()
        fl = oracle.ldltfactor(b, lnz) if spd else oracle.lufactor(b, lnz, unz, ipiv)
        t1 = time.perf_counter()
        X = np.zeros_like(bcols, order="F")
        for j in range(bcols.shape[1]):                 # the reference solves one right-hand side at a time
            rhs = np.ascontiguousarray(bcols[:, j][b.order.rperm - 1])
            if spd:
                oracle.ldltsolve(b, lnz, rhs)
            else:
                oracle.lusolve(b, lnz, unz, ipiv, rhs)
            X[:, j] = rhs[b.order.rinvp - 1]
        t2 = time.perf_counter()
        assert fl == 0
        over = budget_s is not None and time.perf_counter() - t_start > budget_s
        if it >= warmup or over:                         # a warm-up that alone spends the budget is the timed repetition
            if it < warmup:
                warmup = it
            tf.append(t1 - t0); ts.append(t2 - t1)
        it += 1
        if len(tf) >= steps or (over and tf):
            break
    res = rel_residual(A, X.reshape(bb.shape, order="F"), bb)
    return dict(flops=F, factor_s=float(np.mean(tf)), solve_s=float(np.mean(ts)), residual=res, n=b.n, reps=len(tf),
                warmup=warmup, b=b)


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


def dmma_traffic(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "dmma_traffic.json")))
        return d.get(name)
    except Exception:
        return None


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
    ap.add_argument("--config", default=os.environ.get("SPK_BENCH_CONFIG", "cfg4"), choices=sorted(CONFIGS))
    ap.add_argument("--grid", type=int, default=int(os.environ.get("SPK_BENCH_GRID", "0")), help="override the config's grid (tests)")
    ap.add_argument("--cpu-grid", type=int, default=int(os.environ.get("SPK_BENCH_CPU_GRID", "0")))
    ap.add_argument("--ref-budget", type=float, default=float(os.environ.get("SPK_BENCH_REF_BUDGET", "150")),
                    help="reference arm: stop repeating once this many seconds are spent (at least one timed repetition)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the faithful stateless drop-in e2e (lnz both ways)")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel-kind time breakdown to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.config]
    grid = args.grid or cfg["grid"]
    cpu_grid = args.cpu_grid or min(cfg["sample_grid"], grid)
    warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    cores = os.cpu_count()

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
        # the SAME matrix as our arm.  One factorisation of config 4 costs minutes on the host cores, so the
        # repetitions are bounded by --ref-budget and reported (`reps`); the first repetition is timed when a
        # warm-up would not fit.
        r = cpu_reference_run(cfg, grid, max(args.steps, 1), min(args.warmup, 1), budget_s=args.ref_budget)
        gf = r["flops"] / r["factor_s"] / 1e9
        sample = (f"the full workload ({cfg['workload'].format(g=grid)}): oracle restatement of the reference schedule + OpenBLAS, "
                  f"{r['reps']} timed repetition(s) after {r['warmup']} warm-up (bounded by --ref-budget {args.ref_budget:.0f} s)")
        out = {"impl": "reference", "metric": METRIC, "value": gf, "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "reps": r["reps"], "ms_per_step": (r["factor_s"] + r["solve_s"]) * 1e3,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": describe_config(args.config, cfg, grid, r["b"], r["flops"]),
               "factor_s": r["factor_s"], "solve_s": r["solve_s"], "residual": r["residual"],
               "cpu_baseline": {"value": gf, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
               "e2e": {"value": gf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    spk, A, s, dest, nzval = build_problem(cfg, grid)
    b = s.slvr
    nrhs = cfg["nrhs"]
    F, nnzl = structural_flops(b)
    bb = make_rhs(spk, A, nrhs)
    bcols = bb.reshape(b.n, -1, order="F")
    rhs_perm = torch.from_numpy(np.ascontiguousarray(bcols[b.order.rperm - 1, :].T)).cuda()    # (nrhs, n): column-major n x nrhs
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
    log(f"[bench] rank {rank}: fronts={plan.stat(2)} levels={plan.stat(3)} device bytes={plan.stat(4) / 2**30:.1f} GiB "
        f"(arena {plan.stat(6) * 8 / 2**30:.1f}) factor launches={plan.stat(7)} solve launches={plan.stat(8)} "
        f"top-set fronts={plan.stat(13)} structural flops={F:.3e}")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step_resident():
        """one numeric refactorisation + one solve, inputs resident in HBM; returns (flag, factor ms, solve ms)"""
        if ds is None:
            plan.reassemble()                        # arena clear + scatter: inside the factor time (event recorded here)
            fl = plan.factor()
            work.copy_(rhs_perm); torch.cuda.current_stream().synchronize()
            plan.solve_device(work.data_ptr(), nrhs, b.n, 0)
            return fl, plan.statf(2), plan.statf(3)
        sync_all(); t0 = time.perf_counter()
        plan.reassemble()
        fl = ds.factor()                             # subtrees, exchange, top set
        sync_all(); t1 = time.perf_counter()
        work.copy_(rhs_perm)
        ds.solve(work)
        sync_all(); t2 = time.perf_counter()
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
    X = work.cpu().numpy().T[b.order.rinvp - 1, :]
    resid = rel_residual(A, X.reshape(bb.shape, order="F") if nrhs > 1 else X[:, 0], bb)

    # e2e: plan C-ABI with HOST buffers (H2D of nnz(A) values + rhs, D2H of the solution, in the timed region)
    e_ms = []
    nz_host = torch.from_numpy(np.ascontiguousarray(nzval)).pin_memory()     # pinned: what a caller that cares registers once
    xb_host = torch.from_numpy(np.asfortranarray(bcols).T.copy()).pin_memory()    # (nrhs, n) = column-major n x nrhs
    nzval_h = nz_host.numpy()
    xb = xb_host.numpy().T                                                    # Fortran-ordered (n, nrhs) view
    pinned = torch.from_numpy(np.ascontiguousarray(bcols[b.order.rperm - 1, :].T)).pin_memory()
    for it in range(1 + args.steps):
        sync_all(); t1 = time.perf_counter()
        plan.inmatrix(nzval_h)                       # H2D nnz(A) doubles
        if ds is None:
            fl = plan.factor()
            xb[:] = bcols
            plan.triangularsolve(xb if nrhs > 1 else xb[:, 0])       # H2D + D2H n x nrhs doubles
        else:
            fl = ds.factor()
            work.copy_(pinned, non_blocking=False)   # H2D
            ds.solve(work)
            xb[:] = work.cpu().numpy().T[b.order.rinvp - 1, :]       # D2H
        sync_all(); t2 = time.perf_counter()
        if it > 0:
            e_ms.append((t2 - t1) * 1e3)
    e2e_ms = float(np.mean(e_ms))
    e2e_resid = rel_residual(A, np.array(xb) if nrhs > 1 else np.array(xb[:, 0]), bb)

    # roofline of the dominant kernel (DMMA trailing update): per-launch CUDA events on the plan's stream
    plan.stat(100)
    plan.reassemble()
    if ds is None:
        plan.factor()
    else:
        ds.factor()
    gemm_flops, gemm_ms = plan.statf(4), plan.statf(5)
    kinds = ["asm", "asm_tail", "diag", "panel", "gemm_small", "gemm_dmma_128x64", "gemm_dmma_64x64"]
    breakdown = {k: {"ms": plan.statf(10 + i), "launches": int(plan.statf(30 + i))} for i, k in enumerate(kinds)}
    prof_total = plan.statf(2)
    plan.stat(101)
    if args.profile:
        log("[bench] profiled factor (per-launch events, one stream): total %.2f ms" % prof_total)
        for k, v in breakdown.items():
            log(f"    {k:14s} {v['ms']:9.3f} ms  {v['launches']:6d} launches")
    phase_ms = (plan.statf(6), plan.statf(7)) if ds is not None else None
    dev_bytes = plan.stat(4)
    n_sub = plan.stat(15) if world > 1 else 1
    n_dmma_launches = breakdown["gemm_dmma_128x64"]["launches"] + breakdown["gemm_dmma_64x64"]["launches"]
    plan.destroy()
    if ds is not None:
        del ds, eng
    dgemm_tf = measure_dgemm_peak(torch)
    achieved_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0

    # faithful stateless drop-in (`_factor!` contract: lnz overwritten in place => lnz crosses the bus both ways)
    dropin = None
    if world == 1 and not args.no_dropin:
        spk.inmatrix(s)                              # host _inmatrix!: b.lnz / b.unz assembled
        L = _cudalib.lib()
        ipiv = np.zeros(b.n, np.int64)
        t1 = time.perf_counter()
        if b.spd:
            rc = L.spk_ldltfactor_f64(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, b.lnz)
        else:
            rc = L.spk_lufactor_f64(b.n, b.nsuper, b.xsuper, b.snode, b.xlindx, b.lindx, b.xlnz, b.lnz, b.xunz, b.unz, ipiv)
        t2 = time.perf_counter()
        nb = int(b.lnz.nbytes + b.unz.nbytes)
        dropin = {"value": F / (t2 - t1) / 1e9, "unit": UNIT, "ms": (t2 - t1) * 1e3, "rc": int(rc),
                  "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb + (0 if b.spd else int(ipiv.nbytes)),
                  "what": "one call of the stateless _ldltfactor!/_lufactor! drop-in (plan build + pageable lnz/unz H2D + factor + D2H)"}

    tmax = torch.tensor([factor_ms, solve_ms, e2e_ms], dtype=torch.float64, device="cuda")
    rank_phases = None
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ph = torch.tensor([phase_ms[0], phase_ms[1]], dtype=torch.float64, device="cuda")
        allph = [torch.zeros_like(ph) for _ in range(world)]
        dist.all_gather(allph, ph)
        rank_phases = [[round(float(v), 2) for v in t.tolist()] for t in allph]     # per rank: own subtrees, top set (device events)
    factor_ms, solve_ms, e2e_ms = [float(v) for v in tmax.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = F / (factor_ms * 1e-3) / 1e9             # N > 1: ONE factorisation partitioned over the GPUs (strong scaling)
    traffic = dmma_traffic(args.config) if grid == cfg["grid"] and world == 1 else None
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": describe_config(args.config, cfg, grid, b, F),
        "parallelism": (f"{n_sub} elimination subtrees on {world} GPUs, NCCL exchange of update matrices, top separators "
                        f"distributed by column blocks" if world > 1 else "single GPU"),
        "factor_s": factor_ms * 1e-3, "solve_s": solve_ms * 1e-3, "residual": resid, "device_bytes": int(dev_bytes),
        "e2e": {"value": F / (e2e_ms * 1e-3) / 1e9, "unit": UNIT,
                "what": "structural factor flops / wall time of inmatrix + factor + solve through the plan C-ABI, host buffers",
                "ms": e2e_ms, "h2d_bytes_per_step": int(nzval.nbytes + bb.nbytes), "d2h_bytes_per_step": int(bb.nbytes),
                "residual": e2e_resid},
        "e2e_dropin": dropin,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": dgemm_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / dgemm_tf if dgemm_tf > 0 else None,
                     "traffic": traffic["bytes_per_launch"] if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None,
                     "kernel": "k_gemm_dmma (DMMA trailing update)", "peak_source": "cuBLAS DGEMM 8192^3 measured in this run",
                     "kernel_flops": gemm_flops, "kernel_ms": gemm_ms, "launches": n_dmma_launches,
                     "flops_per_launch": gemm_flops / max(n_dmma_launches, 1),
                     "share_of_factor": gemm_ms / prof_total if prof_total else None},
        "breakdown_ms": breakdown, "phase_ms": phase_ms, "rank_phase_ms": rank_phases,
        "wall_s_timed_region": wall,
    }
    if not args.no_cpu_baseline:
        os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
        r = cpu_reference_run(cfg, cpu_grid, 1, 0)
        out["cpu_baseline"] = {"value": r["flops"] / r["factor_s"] / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{cfg['workload'].format(g=cpu_grid)} — bounded sample of the workload, oracle restatement + OpenBLAS "
                                         f"({r['factor_s']:.2f} s factor, {r['solve_s']:.3f} s solve); the reference arm runs the full size"}
    print(json.dumps(out), file=json_out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
