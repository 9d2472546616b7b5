"""Static partition of a unit table across the GPUs of one box.

Units (raw streams, LZMA2 dictionary-reset groups, xz blocks) are independent, so the multi-GPU path is
a partition, not a collective: contiguous ranges of units, balanced by cumulative UNPACKED bytes (the
decode cost driver).  Each rank uploads only the packed bytes of its own range and owns a contiguous
slab of the output.  The only cross-rank traffic is host-side bookkeeping: max-over-ranks time and an
error-code reduction -- done here with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(dst_caps: np.ndarray, world: int) -> np.ndarray:
    """world+1 unit indices; rank r owns units [b[r], b[r+1]).  Balanced by cumulative unpacked bytes.
    The same rule as the C-ABI's Lzb_ShardBounds / Lzb_DecodeUnitsHostMulti (one process, several GPUs); checked
    against it in tests/test_sharding_gloo.py."""
    n = len(dst_caps)
    if world <= 1 or n == 0:
        return np.array([0] + [n] * max(world, 1), dtype=np.int64)
    cum = np.cumsum(dst_caps.astype(np.uint64))
    total = int(cum[-1])
    targets = (np.arange(1, world, dtype=np.float64) * total / world)
    cuts = np.searchsorted(cum, targets, side="left") + 1
    cuts = np.clip(cuts, 0, n)
    b = np.concatenate(([0], cuts, [n])).astype(np.int64)
    return np.maximum.accumulate(b)


def shard_units(units: np.ndarray, world: int, rank: int):
    """This rank's slice of the unit table, rebased to its own packed / output slabs.

    Returns (local_units, src_lo, src_hi, dst_lo, dst_hi, first_index): the rank needs packed bytes
    [src_lo, src_hi) and produces output bytes [dst_lo, dst_hi) of the global arenas.
    """
    b = shard_bounds(units["dst_cap"], world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    loc = units[lo:hi].copy()
    if len(loc) == 0:
        return loc, 0, 0, 0, 0, lo
    src_lo = int(loc["src_off"].min())
    src_hi = int((loc["src_off"] + loc["src_len"]).max())
    dst_lo = int(loc["dst_off"].min())
    dst_hi = int((loc["dst_off"] + loc["dst_cap"]).max())
    loc["src_off"] -= src_lo
    loc["dst_off"] -= dst_lo
    return loc, src_lo, src_hi, dst_lo, dst_hi, lo


def reduce_job(dist, seconds: float, worst_res: int, out_bytes: int, device=None):
    """(max seconds, max SRes, total bytes) over all ranks; plain passthrough when not distributed."""
    if dist is None or not dist.is_initialized():
        return seconds, worst_res, out_bytes
    import torch
    t = torch.tensor([seconds, float(worst_res)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([float(out_bytes)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t[0]), int(t[1]), int(s[0])


def slab_bounds(dist, dst_lo: int, dst_hi: int, device=None):
    """[(dst_lo, dst_hi)] of every rank, in rank order (the slabs tile the output arena)."""
    import torch
    mine = torch.tensor([dst_lo, dst_hi], dtype=torch.int64, device=device)
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    return [tuple(int(v) for v in b.tolist()) for b in every]


def gather_slabs(dist, slab, bounds, full=None, root: int = 0):
    """SURVEY 8e's optional step: collect every rank's output slab at its offset of `full` on `root`.
    Point-to-point (NCCL over NVLink on GPUs, gloo in the CPU test): the slabs are ragged and only one rank wants them,
    so this is one receive per peer posted together, not a collective.  `slab`: this rank's bytes (uint8 tensor of
    exactly bounds[rank] length); `full`: on `root`, a uint8 tensor of bounds[-1][1] bytes."""
    rank = dist.get_rank()
    if rank == root:
        lo, hi = bounds[root]
        full[lo:hi].copy_(slab)
        ops = [dist.P2POp(dist.irecv, full[lo:hi], src) for src, (lo, hi) in enumerate(bounds) if src != root and hi > lo]
    else:
        ops = [dist.P2POp(dist.isend, slab, root)] if slab.numel() else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full
