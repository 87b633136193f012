"""Multi-GPU host logic for the set-abstraction path.

The path shards naturally: every op is independent per cloud (batch entry), exactly the
``tf.slice`` of the feed batch per tower in the reference's train_multi_gpu.py:185-188.  So the
multi-GPU story is one process per GPU, contiguous batch slices, NO data-path collective; the only
cross-rank traffic in a measurement is the max-over-ranks of the timed duration (and, in a full
training step, the outer gradient all-reduce of the MLP weights, which is outside this path).

These helpers are backend-agnostic (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` clouds owned by `rank`; remainders go to the low ranks."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """The rank's slice of a (B, ...) tensor along the batch axis (a view; no copy)."""
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the timed duration): the job is as slow as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job throughput = units processed by ALL ranks / the slowest rank's time."""
    total = sum_over_ranks(units_this_rank, device)
    t = max_over_ranks(seconds_this_rank, device)
    return total / t if t > 0 else float("inf")


def gather_sharded(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather variable-length batch shards back into the full (total, ...) tensor (used by
    tests and by callers that need the full result on every rank; NOT on the measured path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
