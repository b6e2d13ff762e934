"""Multi-GPU plumbing: one process per GPU (torchrun), frames / channel
realisations / SNR points sharded across ranks with no data-path collective.
NCCL (gloo on CPU in the tests) is used for exactly two exchanges (SURVEY.md
§8e): summing error counters, and summing the fp64 Gram matrices when ONE
readout is trained on frames spread over the ranks."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* if the
    process was launched by torchrun; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def world():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def shard_range(n_items, rank_, world_):
    """Contiguous, balanced [begin, end) slice of n_items for this rank (the
    first n_items % world ranks get one extra)."""
    base, extra = divmod(int(n_items), int(world_))
    begin = rank_ * base + min(rank_, extra)
    return begin, begin + base + (1 if rank_ < extra else 0)


def assign_by_cost(costs, world_):
    """Whole work items (sweep configurations) -> ranks, longest-processing-time first: items sorted by
    descending cost, each given to the currently least-loaded rank.  Deterministic (ties by index), so
    every rank computes the same plan without talking.  Returns a list of item-index lists, one per rank,
    each in descending-cost order."""
    loads = [0.0] * int(world_)
    plan = [[] for _ in range(int(world_))]
    for i in sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i)):
        r = min(range(len(loads)), key=lambda r: (loads[r], r))
        plan[r].append(i)
        loads[r] += float(costs[i])
    return plan


def allreduce_sum_(t):
    """In-place sum over ranks (no-op for a single process)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_gram_(G, R):
    """Sum the shared-readout normal equations (G [1,P,P], R [1,P,n_out],
    fp64) over ranks; every rank then runs the same Cholesky."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        flat = torch.cat([G.reshape(-1), R.reshape(-1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        G.copy_(flat[:G.numel()].view_as(G))
        R.copy_(flat[G.numel():].view_as(R))
    return G, R


class GramReducer:
    """Shared-readout training over ranks with the collective hidden behind the harvest: the rank's pilots are
    worked off in chunks; the partial normal equations [G | R] of chunk k (one flat fp64 buffer, 2.2 MB at cfg3) go
    into an ASYNCHRONOUS all-reduce while chunk k + 1 is being harvested; `finish()` waits for the outstanding
    works and adds the reduced partials.  Only the last chunk's all-reduce (tens of microseconds on NVSwitch) is
    exposed.  Single process: no collective, the partials are just added."""

    def __init__(self):
        self.parts, self.works, self.bytes = [], [], 0

    def add(self, flat):
        """flat: this rank's partial [G | R] of one chunk (overwritten by the reduced sum)."""
        self.parts.append(flat)
        if dist.is_initialized() and dist.get_world_size() > 1:
            self.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
            self.bytes += flat.numel() * flat.element_size()

    def finish(self):
        for w in self.works:
            w.wait()
        total = self.parts[0]
        for p in self.parts[1:]:
            total += p
        self.works = []
        return total


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
