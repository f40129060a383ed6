"""Multi-GPU plumbing of the path (one process per GPU, ``torch.distributed``; NCCL on the GPUs, gloo in the CPU
tests).  The path shards over independent units -- test points for posterior evaluation, row blocks for
stand-alone Gram assembly -- so the only exchange step is gathering result rows."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced shard [lo, hi) of ``n`` units for ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_concat(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather 1-D shards of unequal length (``shard_bounds`` order) into the full vector on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros(maxlen, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
