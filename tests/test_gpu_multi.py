"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): CUDA plans, one process per GPU,
NCCL exchange of the subtree-root fronts, against the CPU oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spd, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import sparspak_jl_b200 as spk
        from sparspak_jl_b200.multigpu import CudaEngine, DistributedSolver
        from common import prepare, oracle_factor, spd_mask, rel_err, residual, M
        g = 14
        A = M.laplacian3d(g) if spd else M.convdiff3d(g)
        s = prepare(A, spd, spk.nd_grid_order(g, g, g))
        b = s.slvr
        eng = CudaEngine(b, rank, world, rank)
        eng.plan.set_values(b.lnz, None if spd else b.unz)
        ds = DistributedSolver(eng, rank, world)
        flag = ds.factor()
        ds.gather_factors()
        lo, uo, po, _ = oracle_factor(b)
        lg = np.zeros(b.lnz.size); ug = np.zeros(b.unz.size); pg = np.zeros(b.n, np.int64)
        eng.plan.get_factors(lg, None if spd else ug, None if spd else pg)
        e_l = rel_err(lg, lo, spd_mask(b)); e_u = 0.0 if spd else rel_err(ug, uo)
        piv_ok = True if spd else bool(np.array_equal(pg, po))
        bb = M.rhs_for(A)
        rhs = torch.from_numpy(np.ascontiguousarray(bb[b.order.rperm - 1])).cuda()
        ds.solve(rhs)
        x = rhs.cpu().numpy()[b.order.rinvp - 1]
        q.put((rank, flag, e_l, e_u, piv_ok, residual(A, x, bb)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("spd", [True, False])
def test_two_gpu_subtree_partition(spd):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, spd, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs: p.join(timeout=60)
    for rank, flag, e_l, e_u, piv_ok, resid in res:
        assert flag == 0 and e_l < 1e-11 and e_u < 1e-11 and piv_ok and resid < 1e-12
