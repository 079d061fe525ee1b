"""Sharding of independent units across ranks (one process per GPU, no data-path collective).

BASELINE.json configs[3]: "4096 independent 2k-vertex soft bodies batched and sharded across
1/2/4/8 B200 (no inter-GPU communication)".  Bodies are dealt in contiguous balanced ranges; the
only collectives are the end-of-run reductions of timings and counters (NCCL on GPUs, gloo on CPU).
"""
from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the items owned by `rank`: sizes differ by at most one, lower ranks get the extras."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_max(value: float, device=None) -> float:
    """max over ranks of a host scalar (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
