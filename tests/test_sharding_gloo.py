"""N>1 path on CPU: the static unit partition + the job-level reductions + the optional gather of the output slabs,
world_size 2 and 3 over gloo.
(No decode here -- the units are decoded per rank by the same single-GPU path the gpu tests cover; what
multi-GPU adds is only the partition and the max/sum bookkeeping.)"""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import lzb200
from lzb200 import records as R
from lzb200.sharding import shard_bounds, shard_units


def _table(n, seed):
    rng = np.random.default_rng(seed)
    u = R.make_units(n)
    caps = rng.integers(1, 1 << 20, size=n).astype(np.uint32)
    lens = (caps * 0.6).astype(np.uint32) + 7
    u["dst_cap"], u["src_len"] = caps, lens
    u["dst_off"] = np.concatenate(([0], np.cumsum(caps.astype(np.uint64))[:-1]))
    u["src_off"] = np.concatenate(([0], np.cumsum((lens.astype(np.uint64) + 15) & ~np.uint64(15))[:-1]))
    return u


def test_bounds_cover_and_balance():
    for n, world in ((1000, 2), (1000, 8), (7, 8), (1, 4), (0, 2), (65536, 8)):
        u = _table(n, n + world)
        b = shard_bounds(u["dst_cap"], world)
        assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all() and len(b) == world + 1
        if n >= 64 * world:
            per = np.array([u["dst_cap"][b[r]:b[r + 1]].astype(np.uint64).sum() for r in range(world)], dtype=np.float64)
            assert per.max() / per.mean() < 1.05
    # uniform units split evenly
    u = _table(4096, 1)
    u["dst_cap"] = 65536
    assert np.diff(shard_bounds(u["dst_cap"], 8)).tolist() == [512] * 8


def test_c_abi_shard_bounds_is_the_same_rule():
    """Lzb_ShardBounds (what Lzb_DecodeUnitsHostMulti uses for the GPUs of one process) == sharding.shard_bounds
    (what one-process-per-GPU hosts use).  Host arithmetic only: the library loads without a device."""
    import ctypes as C
    L = lzb200.lib.load()
    L.Lzb_ShardBounds.restype = None
    L.Lzb_ShardBounds.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    for n, world in ((1000, 2), (1000, 8), (7, 8), (1, 4), (0, 2), (65536, 8), (8192, 8), (8192, 3)):
        u = _table(n, 3 * n + world)
        if n == 8192:
            u["dst_cap"] = 1 << 20
        b = np.zeros(world + 1, dtype=np.uint64)
        L.Lzb_ShardBounds(u.ctypes.data, n, world, b.ctypes.data)
        assert b.astype(np.int64).tolist() == shard_bounds(u["dst_cap"], world).tolist(), (n, world)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u = _table(5000, 42)
    loc, src_lo, src_hi, dst_lo, dst_hi, first = shard_units(u, world, rank)
    # every rank "decodes" its shard: here, it just accounts for the bytes it owns
    secs = 0.5 + 0.25 * rank
    worst = 1 if rank == world - 1 else 0            # pretend the last rank met one SZ_ERROR_DATA unit
    t, res, total = lzb200.sharding.reduce_job(dist, secs, worst, int(loc["dst_cap"].astype(np.uint64).sum()))
    import torch
    cover = [None] * world
    dist.all_gather_object(cover, (first, len(loc), src_lo, src_hi, dst_lo, dst_hi))
    # the optional gather of 8e: every rank fills its slab with a function of the GLOBAL byte offset, rank 0 collects
    # (slab sizes scaled down 1024x: the table describes 2.6 GB of output)
    dst_lo, dst_hi = dst_lo >> 10, dst_hi >> 10
    bounds = lzb200.sharding.slab_bounds(dist, dst_lo, dst_hi)
    slab = torch.from_numpy(((np.arange(dst_lo, dst_hi, dtype=np.uint64) * 2654435761 >> 7) & 0xFF).astype(np.uint8))
    full = torch.zeros(bounds[-1][1], dtype=torch.uint8) if rank == 0 else None
    lzb200.sharding.gather_slabs(dist, slab, bounds, full, root=0)
    gathered_ok = None
    if rank == 0:
        want = ((np.arange(0, bounds[-1][1], dtype=np.uint64) * 2654435761 >> 7) & 0xFF).astype(np.uint8)
        gathered_ok = bool(np.array_equal(full.numpy(), want))
    q.put((rank, t, res, total, cover, int(loc["src_off"].min()) if len(loc) else 0, gathered_ok))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    u = _table(5000, 42)
    total = int(u["dst_cap"].astype(np.uint64).sum())
    for rank, t, res, tot, cover, local_min, gathered_ok in got:
        assert gathered_ok is (True if rank == 0 else None)
        assert abs(t - (0.5 + 0.25 * (world - 1))) < 1e-9       # max over ranks
        assert res == 1 and tot == total and local_min == 0
        cover = sorted(cover)
        assert cover[0][0] == 0 and sum(c[1] for c in cover) == 5000
        for a, b in zip(cover, cover[1:]):
            assert a[0] + a[1] == b[0]                            # contiguous unit ranges
            assert a[5] == b[4]                                   # output slabs tile the arena


def test_two_ranks_gloo():
    _run(2)


def test_three_ranks_gloo():
    _run(3)
