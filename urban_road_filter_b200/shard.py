"""Scan-level sharding across ranks: scans are independent units, so a rank only needs to know which scans
are its own; the only exchanges are the barrier, the max-over-ranks time and a sum of per-rank counters. Works on any
torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> range:
    """Contiguous, balanced block of scan indices owned by `rank` (first `total % world` ranks get one extra)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError((total, rank, world))
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def seeds_for_rank(batch: int, rank: int) -> list[int]:
    """Distinct synthetic-scan seeds per rank in the weak-scaling bench (every rank gets `batch` scans of its own)."""
    return [1000 * rank + b for b in range(batch)]


def allreduce_max(values, device=None):
    """Element-wise max over ranks of a list of floats (timings: the job is as slow as its slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def allreduce_sum(values, device=None):
    """Element-wise sum over ranks of a list of integers (label counters used as a cross-rank consistency check)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t]
