"""Multi-GPU numeric factor / solve: one process per GPU, elimination-subtree partition.

Each rank holds a full plan (same structure, `part = rank`): it factors the fronts of its own
subtrees (phase 0), the subtree-ROOT frontal matrices are then broadcast from their owners
(the only exchange of the factorisation: NCCL over NVLink / NVSwitch in production, gloo in the
CPU tests), and every rank factors the small top set of the tree redundantly (phase 1) — bitwise
identical on every rank, so the pivot sequence does not depend on the GPU count.  The solve
mirrors it: forward over the subtrees, broadcast of the subtree-root work vectors, top set,
backward over the subtrees, broadcast of the owned pieces of x.

The orchestration is engine-agnostic: `CudaEngine` drives the CUDA C-ABI plan, the CPU tests
drive the host simulator of the schedule with the same code (tests/test_multigpu_gloo.py)."""
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

    def solve_phase(self, rhs, phase):
        self.plan.solve_phase(rhs.data_ptr(), 1, rhs.numel(), phase)


class DistributedSolver:
    """Factor / solve across `world` ranks.  `engine` provides phases, exchange lists and buffer views."""

    def __init__(self, engine, rank, world):
        self.e, self.rank, self.world = engine, rank, world
        self.fronts = engine.xchg()           # rows: owner, F off, F len, w off, w len, front
        self.rng = engine.ranges()            # rows: owner, lnz off, lnz len, unz off, unz len, col0, ncols

    def factor(self):
        flag = self.e.factor_phase(0)
        F = self.e.F()
        for r in self.fronts:                 # the one exchange step of the factorisation
            dist.broadcast(F[int(r[1]): int(r[1] + r[2])], src=int(r[0]))
        if F.is_cuda:
            torch.cuda.synchronize()
        flag = min(flag, self.e.factor_phase(1))
        t = torch.tensor([flag], dtype=torch.int64, device=F.device)
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
