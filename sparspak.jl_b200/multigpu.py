"""Multi-GPU numeric factor / solve: one process per GPU, elimination-subtree partition.

Each rank holds a plan for `part = rank`: frontal storage for its own subtrees, the top set of the tree and the
subtree roots it receives.  Phase 0 factors the own subtrees; the update matrices of the subtree roots are then
exchanged, and the TOP SET is factored
  * LDL^T: DISTRIBUTED — the columns of every top-set front are dealt to the ranks by outer block; only the owner
    factors a block's panel, broadcasts the factored column slab, and every rank applies the delayed update to
    the column blocks it owns (plan.hpp: build_lists_dist);
  * LU: replicated on every rank (bitwise identical, so the pivot sequence does not depend on the GPU count).
The CUDA engine does all of this inside the C library (`spk_plan_factor_multi` / `spk_plan_solve_multi`, NCCL
broadcasts on the plan's own streams, no host synchronisation between the phases); Python only hands over the
NCCL unique id.  The CPU tests drive the host simulator of the same launch lists over gloo
(tests/test_multigpu_gloo.py): the simulator stops at every broadcast and `DistributedSolver` performs it."""
import numpy as np
import torch
import torch.distributed as dist


class _CudaArray:
    """Zero-copy view of device memory for torch.as_tensor."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def cuda_view(ptr, n, typestr="<f8", device=0):
    if n <= 0 or not ptr:
        return torch.empty(0, dtype=torch.float64 if typestr == "<f8" else torch.int32, device=f"cuda:{device}")
    return torch.as_tensor(_CudaArray(ptr, n, typestr), device=f"cuda:{device}")


class CudaEngine:
    """One CUDA plan (part = rank) + tensor views of its device buffers."""

    def __init__(self, base, rank, world, device):
        from . import _cudalib
        self.plan = _cudalib.Plan(base, device=device, part=rank, nparts=world)
        self.device = device
        self.native_comm = False
        if world > 1 and dist.is_available() and dist.is_initialized():
            # the library owns its NCCL communicator; torch.distributed only carries the 128-byte unique id
            ident = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                ident = torch.from_numpy(self.plan.nccl_unique_id().copy())
            ident = ident.to(f"cuda:{device}") if dist.get_backend() == "nccl" else ident
            dist.broadcast(ident, src=0)
            self.plan.comm_init(ident.cpu().numpy())
            self.native_comm = True
        self.lu = not base.spd
        self.n = int(base.n)
        self._w = None

    def F(self):
        return cuda_view(*self.plan.device_ptr(5), device=self.device)

    def lnz(self):
        return cuda_view(*self.plan.device_ptr(0), device=self.device)

    def unz(self):
        return cuda_view(*self.plan.device_ptr(1), device=self.device)

    def ipiv(self):
        return cuda_view(*self.plan.device_ptr(2), typestr="<i4", device=self.device)

    def w(self):
        return cuda_view(*self.plan.device_ptr(6), device=self.device)

    def xchg(self):
        return self.plan.xchg_list(0)

    def ranges(self):
        return self.plan.xchg_list(1)

    def factor_phase(self, phase):
        return self.plan.factor_phase(phase)

    def factor_multi(self):
        return self.plan.factor_multi()

    def solve_multi(self, rhs):
        nrhs, ld = (1, rhs.numel()) if rhs.dim() == 1 else (rhs.shape[0], rhs.shape[1])     # (nrhs, n) rows = column-major n x nrhs
        self.plan.solve_multi(rhs.data_ptr(), nrhs, ld)

    def solve_phase(self, rhs, phase):
        self.plan.solve_phase(rhs.data_ptr(), 1, rhs.numel(), phase)


class DistributedSolver:
    """Factor / solve across `world` ranks.  `engine` provides phases, exchange lists and buffer views."""

    def __init__(self, engine, rank, world):
        self.e, self.rank, self.world = engine, rank, world
        self.fronts = engine.xchg()           # rows: owner, F off, F len, w off, w len, front
        self.rng = engine.ranges()            # rows: owner, lnz off, lnz len, unz off, unz len, col0, ncols

    def factor(self):
        if getattr(self.e, "native_comm", False):
            flag = self.e.factor_multi()              # phases + exchanges inside the library
        else:
            flag = self.e.factor_phase(0)
            F = self.e.F()
            # the engine runs the top-set list and hands every broadcast (arena offset, length, root) back
            flag = min(flag, self.e.factor_top(lambda ofs, n, root: dist.broadcast(F[ofs: ofs + n], src=root)))
        t = torch.tensor([flag], dtype=torch.int64, device=self.e.F().device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    def gather_factors(self):
        """Make lnz / unz / ipiv complete on every rank (each subtree's storage range comes from its owner)."""
        lnz, unz, ipiv = self.e.lnz(), self.e.unz(), self.e.ipiv()
        for r in self.rng:
            src = int(r[0])
            dist.broadcast(lnz[int(r[1]): int(r[1] + r[2])], src=src)
            if self.e.lu and r[4] > 0:
                dist.broadcast(unz[int(r[3]): int(r[3] + r[4])], src=src)
            if self.e.lu:
                dist.broadcast(ipiv[int(r[5]): int(r[5] + r[6])], src=src)

    def solve(self, rhs):
        """rhs: 1-D tensor (permuted order) on the engine's device; overwritten with the solution on every rank."""
        if getattr(self.e, "native_comm", False):
            self.e.solve_multi(rhs)
            return rhs
        self.e.solve_phase(rhs, 0)
        w = self.e.w()
        for r in self.fronts:
            dist.broadcast(w[int(r[3]): int(r[3] + r[4])], src=int(r[0]))
        if w.is_cuda:
            torch.cuda.synchronize()
        self.e.solve_phase(rhs, 1)
        self.e.solve_phase(rhs, 2)
        for r in self.rng:
            dist.broadcast(rhs[int(r[5]): int(r[5] + r[6])], src=int(r[0]))
        if rhs.is_cuda:
            torch.cuda.synchronize()
        return rhs
