"""Multi-GPU sharding of one read batch: reads are independent, so the batch is cut into contiguous slices balanced
by cumulative bases, one slice per rank (one process per GPU), and the per-read records are gathered on rank 0 in
input order.  There is no data-path collective: torch.distributed only carries the gather of
the small result records and the max-over-ranks timing."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(lengths: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """[begin, end) read indices per rank: contiguous, covering, balanced by cumulative bases."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    if world_size <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world_size, 1) - 1)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        k = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def gather_records(local: np.ndarray, bounds: Sequence[Tuple[int, int]], rank: int, world_size: int, dst: int = 0):
    """Gather the per-read records of all ranks on `dst` in input order (host-side gather; works with gloo and nccl).
    Returns the full array on dst, None elsewhere."""
    import torch.distributed as dist
    if world_size <= 1:
        return local
    parts = [None] * world_size if rank == dst else None
    dist.gather_object(local, parts, dst=dst)
    if rank != dst:
        return None
    n = bounds[-1][1]
    out = np.zeros(n, dtype=local.dtype)
    for r, (b, e) in enumerate(bounds):
        assert len(parts[r]) == e - b, "rank %d returned %d records for a slice of %d" % (r, len(parts[r]), e - b)
        out[b:e] = parts[r]
    return out


def max_over_ranks(value: float, world_size: int) -> float:
    """Device time of a multi-GPU step = max over ranks."""
    if world_size <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
