"""Multi-GPU plumbing: one process per GPU, configs sharded across ranks, ONE exchange step per
CholeskyQR pass (sum of the Gram blocks).  The exchange itself lives in the library (csrc/comm.cu:
ncclAllReduce on the ctx stream inside ipf_solve_gram); torch.distributed only carries the rendezvous
(broadcast of the NCCL unique id) and the CPU (gloo) tests.

The reference has no distributed path (only a commented-out DistributedArrays stub,
`src/lsq_db.jl:202-228,252-265`); its single parallel strategy is the descending-cost
dynamic queue of `tfor` (`src/tools.jl:35-56`), which `shard_configs` mirrors as a
longest-processing-time assignment.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def config_cost(nat: int, nbody_max: int = 3) -> float:
    """Cost model: sites x clusters per site ~ nat * nneigh^(N-1); the neighbour count is
    roughly size independent, so the cost is ~ linear in nat (`src/obsiter.jl:94` uses nat)."""
    return float(nat)


def shard_configs(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """LPT: configs sorted by descending cost (`sortperm(costs, rev=true)`, `src/tools.jl:35`),
    each given to the least loaded rank.  Returns, per rank, the sorted list of config indices."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in out]


class _CudaArrayView:
    """Zero-copy view of a raw device pointer for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


class Communicator:
    """Thin wrapper over torch.distributed (nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.ctx = None           # the ipf context whose library-side communicator mirrors this group

    def attach(self, ctx):
        """Create the library's own NCCL communicator over the ranks of this group (`ipf_comm_init`): from then on
        `ipf_solve_gram` / `ipf_solve_finish` on `ctx` return sums over all ranks and this object only shards."""
        uid = [ctx.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(uid, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                        group=self.group)
        ctx.comm_init(self.world_size, self.rank, uid[0])
        self.ctx = ctx

    def attached(self, ctx=None) -> bool:
        return self.ctx is not None and (ctx is None or ctx is self.ctx)

    def allreduce_device_f64(self, ptr: int, n: int, device: int):
        """In-place sum of n doubles at device address `ptr` across ranks."""
        import torch
        t = torch.as_tensor(_CudaArrayView(ptr, n), device=torch.device("cuda", device))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        torch.cuda.synchronize(device)

    def allreduce_gram(self, g, n: int, device: int):
        """Sum of the per-rank Gram blocks: `g` is a device address (CUDA path) or a numpy array
        (CPU stand-in of the tests), reduced in place."""
        if isinstance(g, np.ndarray):
            g[...] = self.allreduce_numpy(g).reshape(g.shape)
        elif not self.attached():          # attached: ipf_solve_gram has already summed the blocks
            self.allreduce_device_f64(g, n, device)

    def allreduce_numpy(self, a: np.ndarray) -> np.ndarray:
        if self.attached():
            return self.ctx.comm_allreduce(np.asarray(a, dtype=np.float64).ravel()).reshape(np.shape(a))
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allreduce_scalars(self, xs):
        return [float(v) for v in self.allreduce_numpy(np.asarray(xs, dtype=np.float64))]

    def allgather_object(self, obj):
        out = [None] * self.world_size
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def allreduce_named_stats(self, cts, okeys, sse, ssy, cnt):
        """Sum per-(configtype, okey) statistics whose configtype lists differ between ranks."""
        all_cts = []
        for lst in self.allgather_object(list(cts)):
            for ct in lst:
                if ct not in all_cts:
                    all_cts.append(ct)
        ng = 3 * len(all_cts)
        buf = np.zeros((3, max(1, ng)))
        for i, ct in enumerate(cts):
            j = all_cts.index(ct)
            buf[0, 3 * j:3 * j + 3] = sse[3 * i:3 * i + 3]
            buf[1, 3 * j:3 * j + 3] = ssy[3 * i:3 * i + 3]
            buf[2, 3 * j:3 * j + 3] = cnt[3 * i:3 * i + 3]
        buf = self.allreduce_numpy(buf)
        return buf[0], buf[1], np.rint(buf[2]).astype(np.int64), all_cts
