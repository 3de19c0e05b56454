"""Multi-GPU sharding of the consensus path.

Families are independent, so the N-GPU path is a host-side RANGE PARTITION of family indices: rank r
owns the contiguous units [lo_r, hi_r), outputs are concatenated in rank order (the reference
preserves input order, src/lib/reorder_buffer.rs), and the only collective is the end-of-run sum of
the counters (ConsensusCallingStats::merge, caller.rs:278-285).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def partition_by_reads(depths: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous unit ranges balanced by cumulative READ count (not unit count), so a Zipf depth
    distribution still gives every rank the same number of bytes to stream."""
    depths = np.asarray(depths, dtype=np.int64)
    n = int(depths.size)
    if world <= 0:
        raise ValueError("world must be positive")
    cum = np.concatenate([[0], np.cumsum(depths)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        # first unit boundary whose cumulative read count reaches the target
        b = int(np.searchsorted(cum, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def all_reduce_counters(counters, dist=None):
    """Sum a tensor of counters over all ranks (NCCL on GPUs, gloo in the CPU tests)."""
    if dist is None:
        import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters
