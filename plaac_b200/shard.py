"""Process-per-GPU sharding: every rank scores a contiguous, residue-balanced shard of the
same batch and rank 0 receives the records in input order.  Proteins are independent, so the data path has NO
collective; the only communication is the final gather of host-side result records, which works on any
torch.distributed backend (gloo on CPU hosts, nccl where the records are left on the GPUs).

    bounds = shard_bounds(offsets, world)            # identical on every rank (pure arithmetic)
    mine   = score_fn(codes[...], offsets[...])      # this rank's shard -> structured numpy records
    all    = gather_in_order(mine, bounds, rank, world, group)   # rank 0: nprot records; others: None
"""
from __future__ import annotations

import numpy as np

from .capi import shard_plan


def shard_bounds(offsets: np.ndarray, world: int) -> np.ndarray:
    return shard_plan(offsets, world)


def local_slice(codes: np.ndarray, offsets: np.ndarray, bounds: np.ndarray, rank: int):
    """This rank's proteins as (codes, offsets) with offsets rebased to 0."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    o = np.ascontiguousarray(offsets[lo:hi + 1] - offsets[lo], dtype=np.int64)
    c = np.ascontiguousarray(codes[int(offsets[lo]):int(offsets[hi])], dtype=np.uint8)
    return c, o


def gather_in_order(mine: np.ndarray, bounds: np.ndarray, rank: int, world: int, group=None):
    """Gather fixed-size records (any numpy dtype) to rank 0 in input order.  No reduction, no all-to-all."""
    import torch
    import torch.distributed as dist

    nprot = int(bounds[-1])
    itemsize = mine.dtype.itemsize
    counts = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    assert len(mine) == counts[rank], (len(mine), counts[rank])
    if world == 1:
        return mine
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    maxn = max(counts)
    buf = np.zeros(maxn * itemsize, dtype=np.uint8)  # pad to the largest shard: gather needs equal sizes
    buf[:len(mine) * itemsize] = np.frombuffer(mine.tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf).to(device)
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, dst=0, group=group)
    if rank != 0:
        return None
    out = np.zeros(nprot, dtype=mine.dtype)
    for r in range(world):
        lo = int(bounds[r])
        raw = parts[r].cpu().numpy()[:counts[r] * itemsize]
        out[lo:lo + counts[r]] = np.frombuffer(raw.tobytes(), dtype=mine.dtype)
    return out
