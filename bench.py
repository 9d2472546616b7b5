#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the B200 decode path next to the reference's CPU decode.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lzma2_8g|lzma64k|xz4m] [--extras all|none|a,b]
                  [--impl reference]

Headline workload (default) = BASELINE.json configs[2], the configuration the metric is quoted on: ONE LZMA2 stream
of 8 GiB -- 8,192 independent 1 MiB units (dictionary reset each, Lzma2Dec.c:14-26,170-289), a single final EOS --
strong-scaled: the host walker (Lzma2_IndexUnits) cuts the stream at the dictionary resets into ONE unit table,
`sharding.shard_units` gives every rank a contiguous range balanced by unpacked bytes, each rank uploads only its own
packed bytes [src_lo, src_hi) and owns the output slab [dst_lo, dst_hi).  No data-path collective.

  value  : whole-job throughput, device-resident -- every rank's packed slice already in HBM, one fused decode
           launch per rank and step (Lzb_DecodeUnitsDevice), CUDA events on the launching stream, barrier +
           synchronize on both sides, max over ranks; 8 GiB / that time.
  e2e    : the same job through the C-ABI with HOST (pinned) arenas, H2D and D2H inside the timed region
           (Lzb_DecodeUnitsHost on each rank's slice); >= 1 % of the units are then CRC-checked on the host.
  extras : one driver run also yields lines for the other BASELINE configs (key `configs`): config 2 (`weak_lzma64k`:
           65,536 raw 64 KiB streams per GPU, replicated = weak scaling), config 4 (`xz4m`), config 5 (`7z`), config 1
           (`lzma_stream`: one 64 MiB stream, 16 MiB ring dictionary, through LzmaDec_DecodeToBuf 64 KiB in / out).
           Default: all of them at N = 1, `weak_lzma64k` only at N > 1.
  --impl reference : the unmodified reference decoder (oracle/_ref/libref.so, one unit per host thread) on a
           bounded sample of the same workload.  Rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tools.corpus import add_x86_callsites, entropy_sweep_mix, make_corpus  # noqa: E402

METRIC = "decompressed GB/s"
WORKLOADS = {
    "lzma2_8g": "ONE LZMA2 stream of 8 GiB: 8,192 independent 1 MiB units (dict reset each, 1 MiB dictionary, level 1), single "
                "final EOS; unit table from the host walker Lzma2_IndexUnits (BASELINE configs[2])",
    "lzma64k": "65,536 independent 64 KiB raw LZMA streams (LzmaUncompress contract), level 5, lc3 lp0 pb2, dict 64 KiB "
               "(BASELINE configs[1])",
    "xz4m": "xz container blocks of 4 MiB: x86 BCJ + LZMA2 (dict 4 MiB) + CRC64 per block, fused (BASELINE configs[3])",
}
DEFAULT_UNITS = {"lzma2_8g": 8192, "lzma64k": 65536, "xz4m": 2220}   # xz4m: 15 blocks per SM, one per warp
UNIQUE_UNITS = {"lzma2_8g": 256, "lzma64k": 4096, "xz4m": 48}        # >= 192 MiB of uniquely encoded data each


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def log(*a):
    if env_int("RANK", 0) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------ workloads
def build_unique(name: str, rank: int, cache_dir: str, small: bool):
    """The uniquely encoded units of a workload (reference encoder, all host threads).  Rank 0 encodes and caches
    under cache_dir; other ranks load the cache."""
    from oracle import refbind
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"{name}{'_small' if small else ''}_v2.npz")
    if not os.path.exists(path):
        if rank != 0:
            raise RuntimeError("cache missing on non-zero rank")
        ref = refbind.ref()
        thr = host_threads()
        uniq = UNIQUE_UNITS[name] // (16 if small else 1)
        t0 = time.time()
        if name == "lzma64k":
            unit = 65536
            corpus = make_corpus(uniq * unit, seed=1234)
            offs = np.arange(uniq, dtype=np.uint64) * unit
            lens = np.full(uniq, unit, dtype=np.uint32)
            enc = ref.encode_batch_mt(0, corpus, offs, lens, level=5, dict_size=1 << 16, lc=3, lp=0, pb=2, fb=32,
                                      end_mark=0, threads=thr)
            kind = 0
        elif name == "lzma2_8g":
            unit = 1 << 20
            corpus = make_corpus(uniq * unit, seed=1234)
            offs = np.arange(uniq, dtype=np.uint64) * unit
            lens = np.full(uniq, unit, dtype=np.uint32)
            enc = ref.encode_batch_mt(1, corpus, offs, lens, level=1, dict_size=1 << 20, lc=3, lp=0, pb=2, threads=thr)
            kind = 1
        else:
            unit = 4 << 20
            corpus = add_x86_callsites(make_corpus(uniq * unit, seed=1234), seed=7)
            filt = np.empty_like(corpus)
            for k in range(uniq):
                filt[k * unit:(k + 1) * unit] = ref.x86_filter_whole(corpus[k * unit:(k + 1) * unit], 0, 1)
            offs = np.arange(uniq, dtype=np.uint64) * unit
            lens = np.full(uniq, unit, dtype=np.uint32)
            enc = ref.encode_batch_mt(1, filt, offs, lens, level=1, dict_size=4 << 20, lc=3, lp=0, pb=2, threads=thr)
            kind = 2
        crcs = np.array([zlib.crc32(corpus[int(o):int(o) + int(l)].tobytes()) for o, l in zip(offs, lens)], dtype=np.uint64)
        crc64 = np.zeros(uniq, dtype=np.uint64)
        if kind == 2:
            crc64 = np.array([ref.crc64(corpus[int(o):int(o) + int(l)]) for o, l in zip(offs, lens)], dtype=np.uint64)
        blob = np.concatenate([enc["buf"][int(o):int(o) + int(l)] for o, l in zip(enc["off"], enc["len"])])
        np.savez(path + ".tmp.npz", blob=blob, plen=enc["len"], ulen=lens, crc32=crcs, crc64=crc64,
                 props=np.asarray(enc["props"], dtype=np.uint8), kind=kind, encode_secs=enc["secs"])
        os.replace(path + ".tmp.npz", path)
        log(f"{name}: {uniq} unique units encoded in {time.time() - t0:.1f} s")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def global_layout(w, n_units: int, name: str):
    """The whole job as ONE unit table over global arenas (numpy only; nothing is materialised here).
    lzma2_8g: the packed arena IS the LZMA2 stream -- units back to back, one EOS byte at the end."""
    import lzb200
    R = lzb200.records
    uniq = len(w["plen"])
    idx = np.arange(n_units) % uniq
    # lzma2_8g: every unit was encoded as a stream of its own and ends with the 0x00 end-of-stream byte; inside ONE
    # stream only the last unit keeps it, so a unit contributes its bytes without that marker
    trim = 1 if name == "lzma2_8g" else 0
    plen = w["plen"][idx].astype(np.uint64) - np.uint64(trim)
    ulen = w["ulen"][idx].astype(np.uint64)
    step = plen if name == "lzma2_8g" else (plen + 15) & ~np.uint64(15)
    soff = np.concatenate(([0], np.cumsum(step)[:-1])).astype(np.uint64)
    doff = np.concatenate(([0], np.cumsum(ulen)[:-1])).astype(np.uint64)
    ustart = np.concatenate(([0], np.cumsum(w["plen"].astype(np.uint64))[:-1]))
    units = R.make_units(n_units)
    units["src_off"], units["src_len"], units["dst_off"], units["dst_cap"] = soff, plen, doff, ulen
    kind = int(w["kind"])
    if kind == 0:
        lc, lp, pb, ds = R.parse_lzma_props(w["props"])
        units["dict_size"], units["codec"], units["lc"], units["lp"], units["pb"] = ds, R.CODEC_LZMA, lc, lp, pb
        units["finish"] = R.LZMA_FINISH_ANY
    else:
        prop = int(w["props"][0])
        units["dict_size"] = (2 | (prop & 1)) << (prop // 2 + 11)
        units["codec"], units["lc"], units["pb"], units["finish"] = R.CODEC_LZMA2, 3, 2, R.LZMA_FINISH_END
        if kind == 2:
            units["filter"], units["check"], units["flags"] = R.FILTER_X86, R.CHECK_CRC64, R.UNIT_VERIFY
            units["expect"] = w["crc64"][idx]
    return dict(idx=idx, units=units, ustart=ustart, trim=trim, stream_bytes=int(soff[-1] + step[-1]) + (1 if name == "lzma2_8g" else 0),
                sum_p=int(plen.sum()), sum_u=int(ulen.sum()))


def materialise(dst: np.ndarray, w, g, first: int, count: int, src_lo: int):
    """Packed bytes of units [first, first + count) of the global job at their place relative to src_lo."""
    blob, gu = w["blob"], g["units"]
    for i in range(first, first + count):
        k = int(g["idx"][i])
        s, n = int(g["ustart"][k]), int(w["plen"][k]) - g["trim"]
        o = int(gu["src_off"][i]) - src_lo
        dst[o:o + n] = blob[s:s + n]
    if g["trim"] and first + count == len(gu):
        dst[g["stream_bytes"] - 1 - src_lo] = 0                      # the stream's single end-of-stream byte


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------ CPU legs
def cpu_reference_run(w, g, name: str, threads: int, sample_units: int, repeat: int = 1):
    """The unmodified reference decoder, one unit per host thread pulled from a shared counter."""
    from oracle import refbind
    ref = refbind.ref()
    gu = g["units"]
    n = min(sample_units, len(gu))
    kind = int(w["kind"])
    best = None
    soff = gu["src_off"][:n].astype(np.uint64)
    plen = gu["src_len"][:n].astype(np.uint32)
    doff = gu["dst_off"][:n].astype(np.uint64)
    ulen = gu["dst_cap"][:n].astype(np.uint32)
    packed = np.zeros(int(soff[n - 1] + plen[n - 1]) + 64, dtype=np.uint8)
    materialise(packed, w, g, 0, n, 0)
    if kind == 2:
        # xz leg: each block wrapped as a one-block .xz member so the whole XzUnpacker path (block header,
        # LZMA2, BCJ, CRC-64, index, footer) is timed, one member per host thread
        uniq = len(w["plen"])
        ustart = g["ustart"]
        members = [ref.xz_assemble(w["blob"][int(ustart[k]):int(ustart[k]) + int(w["plen"][k])], [0], [int(w["plen"][k])],
                                   [int(w["ulen"][k])], [int(w["crc64"][k])], 4, 4, 0, int(w["props"][0])) for k in range(min(uniq, n))]
        idx = g["idx"][:n]
        mlen = np.array([len(members[int(k)]) for k in idx], dtype=np.uint32)
        moff = np.concatenate(([0], np.cumsum((mlen.astype(np.uint64) + 15) & ~np.uint64(15))[:-1])).astype(np.uint64)
        arena = np.zeros(int(moff[-1]) + int(mlen[-1]) + 64, dtype=np.uint8)
        for i, k in enumerate(idx):
            arena[int(moff[i]):int(moff[i]) + int(mlen[i])] = members[int(k)]
        for _ in range(repeat):
            r = ref.xz_decode_batch_mt(arena, moff, mlen, doff, ulen, threads, scratch=True)
            assert (r["res"] == 0).all() and (r["dst_done"] == ulen).all(), "reference xz decode failed"
            best = r["secs"] if best is None else min(best, r["secs"])
        return dict(secs=best, out_bytes=int(ulen.astype(np.uint64).sum()), units=n)
    for _ in range(repeat):
        r = ref.decode_batch_mt(0 if kind == 0 else 1, packed, soff, plen, doff, ulen, w["props"],
                                0 if kind == 0 else 1, threads, scratch=True)
        assert (r["res"] == 0).all() and (r["dst_done"] == ulen).all(), "reference decode failed"
        best = r["secs"] if best is None else min(best, r["secs"])
    return dict(secs=best, out_bytes=int(ulen.astype(np.uint64).sum()), units=n)


def cpu_sample_units(name: str, threads: int, n_units: int) -> int:
    per_thread = {"lzma64k": 512, "lzma2_8g": 16, "xz4m": 4}[name]      # ~10-20 s of CPU work on all host threads
    return min(n_units, max(2 * threads, per_thread * threads))


# ------------------------------------------------------------------------------------ one unit workload on the GPUs
class Job:
    """One unit workload, sharded (strong) or replicated (weak) over the ranks."""

    def __init__(self, name, n_units, a, env, replicate=False):
        import lzb200
        self.name, self.a, self.env, self.replicate = name, a, env, replicate
        rank, world, dist = env["rank"], env["world"], env["dist"]
        if rank == 0:
            self.w = build_unique(name, 0, env["cache"], a.small)
        if dist:
            dist.barrier()
        if rank != 0:
            self.w = build_unique(name, rank, env["cache"], a.small)
        self.g = global_layout(self.w, n_units, name)
        gu = self.g["units"]
        if replicate or world == 1:
            self.first, self.count, self.src_lo, self.dst_lo = 0, n_units, 0, 0
            self.src_hi, self.dst_hi = self.g["stream_bytes"], self.g["sum_u"]
            self.local = gu.copy()
        else:
            self.local, self.src_lo, self.src_hi, self.dst_lo, self.dst_hi, self.first = lzb200.sharding.shard_units(gu, world, rank)
            self.count = len(self.local)
            if name == "lzma2_8g" and rank == world - 1:
                self.src_hi += 1                                                 # the stream's EOS byte
        self.n_global = n_units

    def stage(self):
        """Pinned host arenas for this rank's slice; for lzma2_8g the unit table the GPU gets is the HOST WALKER's
        (Lzma2_IndexUnits over the materialised slice of the stream), checked against the layout arithmetic."""
        import lzb200
        S = lzb200.sdk
        self.h_packed = S.pinned_empty(self.src_hi - self.src_lo + 64)
        self.h_packed[self.src_hi - self.src_lo:] = 0
        materialise(self.h_packed, self.w, self.g, self.first, self.count, self.src_lo)
        self.h_out = S.pinned_empty(self.dst_hi - self.dst_lo + 64)
        self.walker = None
        if self.name == "lzma2_8g":
            t0 = time.perf_counter()
            ix = S.lzma2_index_units(self.h_packed[: self.src_hi - self.src_lo], int(self.w["props"][0]), max_units=self.count + 8)
            dt = time.perf_counter() - t0
            wu = ix["units"]
            same = ix["res"] == 0 and ix["n"] == self.count and all(
                np.array_equal(wu[f], self.local[f]) for f in ("src_off", "src_len", "dst_off", "dst_cap", "dict_size", "codec"))
            if not same:
                raise SystemExit(f"bench.py: Lzma2_IndexUnits disagrees with the layout of the stream (res {ix['res']}, n {ix['n']} vs {self.count})")
            self.local = wu                     # finish = the walker's "unit" mode: the last chunk's termination is validated
            self.walker = dict(api="Lzma2_IndexUnits", units=int(ix["n"]), ms=dt * 1e3, has_eos=bool(ix["has_eos"]))

    def run(self, steps, warmup, e2e=True, verify_frac=0.01, gather=False):
        import torch
        import lzb200
        R = lzb200.records
        env = self.env
        dev, ctx, dist, world = env["dev"], env["ctx"], env["dist"], env["world"]
        units = self.local
        n = self.count
        # measurement aid (profiles/r02_z_x86_tiles.md): LZB_BENCH_POST=none|crc|x86 strips the fused post pass (or half of
        # it) from an xz4m unit table, so that the difference in kernel time is what it costs; such a line is not verified
        post_mode = os.environ.get("LZB_BENCH_POST", "all") if int(self.w["kind"]) == 2 else "all"
        if post_mode != "all":
            units = units.copy()
            if post_mode in ("none", "crc"):
                units["filter"] = R.FILTER_NONE
            if post_mode in ("none", "x86"):
                units["check"], units["flags"], units["expect"] = R.CHECK_NONE, 0, 0
        max_lclp, max_pb = int((units["lc"].astype(int) + units["lp"]).max()), int(units["pb"].max())
        pbytes = self.src_hi - self.src_lo
        d_packed = torch.empty(pbytes + 64, dtype=torch.uint8, device=dev)
        d_packed.copy_(torch.from_numpy(self.h_packed), non_blocking=False)
        d_out = torch.empty(self.dst_hi - self.dst_lo + 64, dtype=torch.uint8, device=dev)
        d_units = torch.from_numpy(units.view(np.uint8).reshape(-1)).to(dev)
        d_results = torch.zeros(n * R.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        stream = torch.cuda.Stream(device=dev)      # the stream the decode kernels are launched on AND timed on
        torch.cuda.synchronize()

        def step_device(u=d_units):
            lzb200.sdk.decode_units_device(ctx, d_packed.data_ptr(), d_out.data_ptr(), u.data_ptr(), d_results.data_ptr(), n,
                                           max_lclp, max_pb, stream.cuda_stream)

        def sync_all():
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
                torch.cuda.synchronize()

        def reduce_max(x: float) -> float:
            if not dist:
                return x
            tt = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        def reduce_all(flag: bool) -> bool:
            if not dist:
                return flag
            tt = torch.tensor([0.0 if flag else 1.0], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()) == 0.0

        for _ in range(warmup):
            step_device()
        sync_all()
        sampler = ClockSampler(env["local"])
        sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for k in range(steps):
            step_device()
            ev[k + 1].record(stream)
        sync_all()
        clocks = sampler.stop()
        my_ms = ev[0].elapsed_time(ev[-1])
        total_ms = reduce_max(my_ms)
        kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        ms_per_step = total_ms / steps
        job_bytes = self.g["sum_u"] * (world if self.replicate else 1)
        value = job_bytes / (ms_per_step * 1e-3) / 1e9

        # correctness of what was just timed (outside the timed region): statuses + every unit's CRC on the device
        res = d_results.cpu().numpy().view(R.RESULT_DTYPE)
        ok = bool((res["res"] == 0).all() and (res["dst_done"] == units["dst_cap"]).all())
        vunits = units.copy()
        lidx = self.g["idx"][self.first:self.first + n]
        if int(self.w["kind"]) != 2:
            vunits["check"], vunits["flags"], vunits["expect"] = R.CHECK_CRC32, R.UNIT_VERIFY, self.w["crc32"][lidx]
        d_vunits = torch.from_numpy(vunits.view(np.uint8).reshape(-1)).to(dev)
        step_device(d_vunits)
        torch.cuda.synchronize()
        vres = d_results.cpu().numpy().view(R.RESULT_DTYPE)
        verified = reduce_all(ok and bool((vres["res"] == 0).all())) if post_mode == "all" else None
        del d_vunits
        gather_d = self.gather_on_rank0(d_out, verify_frac) if (gather and dist and not self.replicate) else None

        # e2e: host arenas through the C-ABI, copies inside the timed region
        e2e_d = None
        if e2e:
            del d_out, d_packed                     # the C-ABI call owns its own device arenas
            torch.cuda.empty_cache()
            for _ in range(max(min(warmup, 2), 1)):
                lzb200.sdk.decode_units_host(ctx, self.h_packed, self.h_out, units)
            sync_all()
            t0 = time.perf_counter()
            per_call = []
            for _ in range(steps):
                tc = time.perf_counter()
                hres = lzb200.sdk.decode_units_host(ctx, self.h_packed, self.h_out, units)
                per_call.append((time.perf_counter() - tc) * 1e3)
            torch.cuda.synchronize()
            dt = reduce_max(time.perf_counter() - t0)
            if dist:
                dist.barrier()
            # host-side check of the bytes the caller received: zlib.crc32 of >= 1 % of the units (evenly spread)
            e2e_ok = bool((hres["res"] == 0).all())
            stride = max(1, int(1.0 / max(verify_frac, 1e-6)))
            checked = 0
            for j in range(0, n, stride):
                o, ln = int(units["dst_off"][j]), int(units["dst_cap"][j])
                e2e_ok = e2e_ok and zlib.crc32(self.h_out[o:o + ln].tobytes()) == int(self.w["crc32"][lidx[j]])
                checked += 1
            e2e_ok = reduce_all(e2e_ok)
            e2e_d = dict(value=job_bytes / (dt / steps) / 1e9, unit="GB/s",
                         h2d_bytes_per_step=int(units["src_len"].astype(np.uint64).sum() + units.nbytes),
                         d2h_bytes_per_step=int(units["dst_cap"].astype(np.uint64).sum() + res.nbytes),
                         ms_per_step=dt / steps * 1e3, per_call_ms=[round(x, 2) for x in per_call], verified=e2e_ok,
                         host_crc32_checked_units=checked, bytes_are="this rank's" if world > 1 else "the job's",
                         api="Lzb_DecodeUnitsHost (pinned host arenas; one launch fed and drained while it runs)")
        kms = float(np.mean(kernel_ms))
        return dict(value=value, ms_per_step=ms_per_step, kernel_ms=kms, clocks=clocks, verified=verified, e2e=e2e_d,
                    alg_bytes=int(units["src_len"].astype(np.uint64).sum() + units["dst_cap"].astype(np.uint64).sum()),
                    units_this_rank=n, my_ms_per_step=my_ms / steps, gather=gather_d)

    def gather_on_rank0(self, d_out, verify_frac):
        """SURVEY 8e's optional step: the ranks' output slabs are contiguous ranges of the decoded stream, so rank 0
        receives each one at its offset (NCCL point-to-point over NVLink) and ends up with the whole stream on one GPU."""
        import torch
        env = self.env
        dev, dist, world, rank = env["dev"], env["dist"], env["world"], env["rank"]
        import lzb200
        bounds = lzb200.sharding.slab_bounds(dist, self.dst_lo, self.dst_hi, dev)
        total = bounds[-1][1]
        d_full = torch.empty(total, dtype=torch.uint8, device=dev) if rank == 0 else None
        slab = d_out[: self.dst_hi - self.dst_lo]
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lzb200.sharding.gather_slabs(dist, slab, bounds, d_full, root=0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            best = float(tt.item()) if best is None else min(best, float(tt.item()))
        if rank != 0:
            return None
        gu, ok, checked = self.g["units"], True, 0
        stride = max(1, int(1.0 / max(verify_frac, 1e-6)))
        for j in range(0, len(gu), stride):
            o, ln = int(gu["dst_off"][j]), int(gu["dst_cap"][j])
            ok = ok and zlib.crc32(d_full[o:o + ln].cpu().numpy().tobytes()) == int(self.w["crc32"][self.g["idx"][j]])
            checked += 1
        return dict(to="rank 0's GPU", api="torch.distributed batch_isend_irecv (NCCL point-to-point over NVLink)", bytes=total,
                    bytes_received=total - (bounds[0][1] - bounds[0][0]), ms=best, GBps=total / (best * 1e-3) / 1e9,
                    verified=bool(ok), host_crc32_checked_units=checked,
                    note="outside the timed region; best of 3; the device-to-device copy of rank 0's own slab is included")


def roofline_of(name: str, r: dict, geometry_note: str):
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = r["alg_bytes"] / (r["kernel_ms"] * 1e-3) / 1e9
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            rec = json.load(open(tp)).get(name)
            if rec:   # measured DRAM bytes per algorithmic byte (ncu dram__bytes_read/write.sum of this launch shape)
                traffic = int(rec["dram_bytes_per_algorithmic_byte"] * r["alg_bytes"])
                tsrc = rec.get("source")
        except Exception:
            traffic = None
    return dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=tsrc,
                kernel="lzb_decode_units_kernel<shared model, G lanes per unit> (rank 0's launch)", launches_per_step=1,
                kernel_ms=r["kernel_ms"], algorithmic_bytes_per_launch=r["alg_bytes"], peak_source=peak_src,
                note="serial range-decoder chain: bound by resident units x issue slots (and, with one unit per warp, by the "
                     "unit's own dependency chain), far below the HBM roof. " + geometry_note)


def geometry_note(n_units_rank: int, sms: int) -> str:
    """What lzb_plan_few_units picks at the default properties (36 resident units per SM): DESIGN.md 2 / 4a."""
    per_sm = n_units_rank / max(sms, 1)
    if n_units_rank <= sms * 16:
        return (f"{n_units_rank} units on this GPU = {per_sm:.1f} per SM: one unit per warp in one wave, look-ahead form of the decoder "
                "(lzb_decode_units_kernel<1,32,*,1>); the limiter is the issue rate of a unit's own warp, not a collective")
    if n_units_rank <= sms * 29:
        return (f"{n_units_rank} units on this GPU = {per_sm:.1f} per SM: one unit per warp, 16 persistent warps per SM taking units from "
                "the ticket, look-ahead form (lzb_decode_units_kernel<1,32,*,1>)")
    return (f"{n_units_rank} units on this GPU = {per_sm:.1f} per SM over 36 resident (two per warp, lean form, "
            "lzb_decode_units_kernel<1,16,*,0>): issue-slot bound")


# ------------------------------------------------------------------------------------ extras: configs 5 and 1
def extra_7z(env, steps=2):
    """BASELINE config 5 through the container API: SzArchive_Open + SzArchive_ExtractAll on a 7z archive of many
    single-file LZMA folders (entropy sweep from byte runs to random), folder CRC-32 fused, packed header."""
    import lzb200
    from oracle import refbind
    from tools import sz_writer as W
    ref = refbind.ref()
    thr = host_threads()
    n_folders, file_bytes = 32768, 65536
    parts = [make_corpus(n_folders // 16 * file_bytes, seed=500 + k, mix=entropy_sweep_mix(k / 15.0)) for k in range(16)]
    corpus = np.concatenate(parts)
    n = len(corpus) // file_bytes
    offs = np.arange(n, dtype=np.uint64) * file_bytes
    lens = np.full(n, file_bytes, dtype=np.uint32)
    enc = ref.encode_batch_mt(0, corpus, offs, lens, level=1, dict_size=1 << 16, lc=3, lp=0, pb=2, fb=32, end_mark=0, threads=thr)

    def archive_of(lo, hi):
        folders = []
        for i in range(lo, hi):
            data = corpus[i * file_bytes:(i + 1) * file_bytes].tobytes()
            f = W.Folder("lzma", [(f"d{i % 97:02d}/f{i:06d}.bin", data)])
            f.packed = enc["buf"][int(enc["off"][i]):int(enc["off"][i]) + int(enc["len"][i])].tobytes()
            f.props = enc["props"].tobytes()
            f.plain = data
            folders.append(f)
        return W.write_archive(ref, folders, encode_header=True)

    arc = np.frombuffer(archive_of(0, n), dtype=np.uint8)
    L = lzb200.lib.load()
    L.SzArchive_Open.restype = C.c_int
    L.SzArchive_Open.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.SzArchive_Free.argtypes = [C.c_void_p]
    L.SzArchive_ExtractAll.restype = C.c_int
    L.SzArchive_ExtractAll.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    out = lzb200.sdk.pinned_empty(len(corpus) + 64)
    harc = lzb200.sdk.pinned_empty(len(arc)); harc[:] = arc
    fres, ires = np.zeros(n + 8, np.int32), np.zeros(n + 8, np.int32)
    times, open_ms = [], []
    for i in range(2 + steps):
        t0 = time.perf_counter()
        h = C.c_void_p(None)
        r1 = L.SzArchive_Open(env["ctx"].h, harc.ctypes.data, len(harc), C.byref(h))
        t1 = time.perf_counter()
        r2 = L.SzArchive_ExtractAll(env["ctx"].h, h, harc.ctypes.data, len(harc), out.ctypes.data, len(corpus), fres.ctypes.data, ires.ctypes.data)
        L.SzArchive_Free(h)
        if i >= 2:
            times.append(time.perf_counter() - t0)
            open_ms.append((t1 - t0) * 1e3)
        assert r1 == 0 and r2 == 0 and not fres[:n].any() and not ires[:n].any()
    ok = zlib.crc32(out[: len(corpus)].tobytes()) == zlib.crc32(corpus.tobytes())
    # the reference: one thread per archive (SzArEx_Extract loop), bounded sample = an archive of 512 folders spread
    # over the entropy sweep
    pick = np.linspace(0, n - 1, 512).astype(int)
    folders = []
    for i in pick:
        data = corpus[i * file_bytes:(i + 1) * file_bytes].tobytes()
        f = W.Folder("lzma", [(f"f{i:06d}.bin", data)])
        f.packed = enc["buf"][int(enc["off"][i]):int(enc["off"][i]) + int(enc["len"][i])].tobytes()
        f.props = enc["props"].tobytes()
        f.plain = data
        folders.append(f)
    sub = W.write_archive(ref, folders, encode_header=True)
    t0 = time.perf_counter()
    r = ref.sz_extract_all(sub, out_cap=len(pick) * file_bytes + 4096, max_files=len(pick) + 8)
    cpu = time.perf_counter() - t0
    assert r["open_res"] == 0 and all(f["res"] == 0 for f in r["files"])
    total = int(len(corpus))
    lzb200.sdk.pinned_free_all()
    return dict(workload=f"7z archive of {n} single-file LZMA folders of {file_bytes >> 10} KiB, entropy sweep byte-runs..random, packed "
                "header, folder CRC-32 fused (BASELINE configs[4])", api="SzArchive_Open + SzArchive_ExtractAll + SzArchive_Free (pinned host buffers)",
                value=total / float(np.mean(times)) / 1e9, unit="GB/s", e2e=True, ms_per_call=[round(x * 1e3, 1) for x in times],
                open_ms=[round(x, 1) for x in open_ms], archive_bytes=int(len(arc)), out_bytes=total, verified=bool(ok),
                cpu_baseline=dict(value=len(pick) * file_bytes / cpu / 1e9, unit="GB/s", cores=1, kind="reference",
                                  sample=f"{len(pick)} folders spread over the sweep; SzArEx_Open + SzArEx_Extract per file, one thread "
                                         "(how the reference extracts an archive)"))


def extra_lzma_stream(env):
    """BASELINE config 1: ONE 64 MiB raw LZMA stream, lc3 lp0 pb2, 16 MiB dictionary, decoded like LzmaUtil.c:55-106:
    LzmaDec_Allocate + a loop over LzmaDec_DecodeToBuf with 64 KiB in / 64 KiB out.  One stream is one serial chain
    (no parallelism to give a GPU): replicas only; the number is reported, it is not a target."""
    import lzb200
    from oracle import refbind
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from streamdrive import GpuStream
    ref = refbind.ref()
    size = (4 << 20) if env["small"] else (64 << 20)
    plain = make_corpus(size, seed=4321)
    level = 1                     # the match finder of level 5 needs ~40 s for 64 MiB on one thread; the decoder's work per
    t0 = time.perf_counter()      # byte is the same order (dictionary, lc/lp/pb as LzmaUtil's defaults)
    packed, props = ref.lzma_encode(plain, level=level, dict_size=16 << 20, end_mark=1)
    enc_s = time.perf_counter() - t0
    g = GpuStream()
    t0 = time.perf_counter()
    b = g.lzma_decode_tobuf(packed, props, len(plain), 1 << 16, 1 << 16, 0)
    gpu_s = time.perf_counter() - t0
    ok = b["res"] == 0 and b["dest_len"] == len(plain) and zlib.crc32(b["out"][: len(plain)].tobytes()) == zlib.crc32(plain.tobytes())
    t0 = time.perf_counter()
    a = ref.lzma_decode_tobuf(packed, props, len(plain), 1 << 16, 1 << 16, 0)
    cpu_s = time.perf_counter() - t0
    calls = (len(plain) + 65535) // 65536 + (len(packed) + 65535) // 65536
    return dict(workload=f"one {size >> 20} MiB raw LZMA stream, lc3 lp0 pb2, 16 MiB dictionary (ring), level {level}, end mark "
                "(BASELINE configs[0])", api="LzmaDec_Allocate + LzmaDec_DecodeToBuf loop, 64 KiB in / 64 KiB out (LzmaUtil.c:55-106)",
                value=len(plain) / gpu_s / 1e9, unit="GB/s", e2e=True, seconds=gpu_s, calls_about=int(calls),
                ms_per_call=gpu_s * 1e3 / max(calls, 1), verified=bool(ok and a["res"] == 0), packed_bytes=int(len(packed)), encode_seconds=enc_s,
                cpu_baseline=dict(value=len(plain) / cpu_s / 1e9, unit="GB/s", cores=1, kind="reference", sample="the same stream, same loop"),
                note="a single range-coded stream is one dependency chain: one warp of one SM works, replicas only")


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lzma2_8g", choices=sorted(WORKLOADS) + ["lzma2_1m"])
    ap.add_argument("--units", type=int, default=0, help="units of the whole job (default: the BASELINE size of the workload)")
    ap.add_argument("--small", action="store_true", help="tiny corpus, for smoke-testing the harness")
    ap.add_argument("--weak", action="store_true", help="replicate the job on every rank instead of sharding it")
    ap.add_argument("--extras", default="auto", help="auto | all | none | comma list of weak_lzma64k,xz4m,7z,lzma_stream")
    ap.add_argument("--gather", action="store_true", help="N > 1: after the timed steps, collect every rank's output slab on rank 0's GPU "
                                                          "over NVLink (NCCL send/recv) and report the time as an extra key")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    if a.workload == "lzma2_1m":
        a.workload = "lzma2_8g"

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    cache = os.environ.get("LZB_BENCH_CACHE", "/tmp/lzb_bench_cache")
    n_units = a.units or (DEFAULT_UNITS[a.workload] // 16 if a.small else DEFAULT_UNITS[a.workload])
    threads = host_threads()
    weak = a.weak
    uniq = UNIQUE_UNITS[a.workload] // (16 if a.small else 1)
    cfg = dict(workload=WORKLOADS[a.workload], units=n_units, unique_units=min(uniq, n_units),
               sharding=("replicated: every rank decodes the whole table" if weak else
                         "static: ONE unit table, contiguous ranges balanced by unpacked bytes (sharding.shard_units); each rank "
                         "uploads only its packed slice and owns its output slab; no collective"),
               l2="inputs >> L2 (126 MB): every tile is a physical copy",
               corpus="synthetic mixed-entropy, seed 1234, reference encoder; the job is tiled from `unique_units` uniquely "
                      "encoded units (copies of one unit never share a warp)")

    # ---------------- reference arm: CPU only, rank 0 only
    if a.impl == "reference":
        if rank != 0:
            return 0
        w = build_unique(a.workload, 0, cache, a.small)
        g = global_layout(w, n_units, a.workload)
        sample = cpu_sample_units(a.workload, threads, n_units)
        if a.warmup > 0:
            cpu_reference_run(w, g, a.workload, threads, min(sample, 2 * threads))
        times = [cpu_reference_run(w, g, a.workload, threads, sample) for _ in range(a.steps)]
        secs = float(np.mean([x["secs"] for x in times]))
        v = times[0]["out_bytes"] / secs / 1e9
        cfg["sample_units"] = times[0]["units"]
        line = dict(metric=METRIC, value=v, unit="GB/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=secs * 1e3, higher_is_better=True, scaling="weak" if weak else "strong", vs_baseline=None, dtype="u8",
                    data="synthetic", config=cfg, impl="reference",
                    cpu_baseline=dict(value=v, unit="GB/s", cores=threads, kind="reference",
                                      sample=f"{times[0]['units']} units ({times[0]['out_bytes'] / 2**20:.0f} MiB out) per step -- a bounded "
                                             f"sample of the job, so ms_per_step is the sample's time, not the job's -- one unit per thread, "
                                             f"unmodified LzmaDec/Lzma2Dec/XzUnpacker (oracle/_ref/libref.so)"),
                    e2e=dict(value=v, unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    # ---------------- B200 arm
    import torch
    import lzb200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # NCCL prints its version banner on stdout when the first communicator comes up; stdout carries the ONE JSON
        # line of the contract, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    ctx = lzb200.lib.Context(local)
    env = dict(rank=rank, world=world, local=local, dist=dist, cache=cache, dev=torch.device("cuda", local), ctx=ctx, small=a.small)

    t_start = time.time()
    job = Job(a.workload, n_units, a, env, replicate=weak)
    job.stage()
    log(f"{a.workload}: staged {job.count} of {n_units} units on rank 0 in {time.time() - t_start:.1f} s")
    r = job.run(a.steps, a.warmup, e2e=not a.no_e2e, gather=a.gather)
    log(f"{a.workload}: value {r['value']:.3f} GB/s, e2e {r['e2e']['value'] if r['e2e'] else None}")
    note = geometry_note(r["units_this_rank"], ctx.sms or 148)
    roofline = roofline_of(a.workload, r, note)
    cfg["units_per_gpu"] = r["units_this_rank"]
    if job.walker:
        cfg["unit_table"] = job.walker

    # whole-stream container call (N = 1 only): the reference-facing API on the 8 GiB stream in host memory
    container = None
    if a.workload == "lzma2_8g" and world == 1 and not a.no_e2e:
        L = lzb200.lib.load()
        total = job.g["sum_u"]
        times = []
        for i in range(1 + min(a.steps, 2)):
            dl, sl, st = C.c_size_t(total), C.c_size_t(job.g["stream_bytes"]), C.c_int(0)
            t0 = time.perf_counter()
            rr = L.Lzma2DecodeParallel(ctx.h, job.h_out.ctypes.data, C.byref(dl), job.h_packed.ctypes.data, C.byref(sl),
                                       int(job.w["props"][0]), 1, C.byref(st))
            if i >= 1:
                times.append(time.perf_counter() - t0)
        okc = rr == 0 and st.value == 1 and dl.value == total and sl.value == job.g["stream_bytes"] and \
            zlib.crc32(job.h_out[: 1 << 20].tobytes()) == int(job.w["crc32"][0])
        container = dict(api="Lzma2DecodeParallel = Lzma2Dec_DecodeToDic contract on the whole stream (host buffers; chunk-header walk on "
                             "the host, all units in one GPU pass)", value=total / float(np.mean(times)) / 1e9, unit="GB/s",
                         res=int(rr), status=int(st.value), dest_len=int(dl.value), src_len=int(sl.value), verified=bool(okc))

    # ---------------- CPU baseline on this box (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import refbind
        if refbind.have_ref():
            sample = cpu_sample_units(a.workload, threads, n_units)
            cpu_reference_run(job.w, job.g, a.workload, threads, min(sample, 2 * threads))
            rc = cpu_reference_run(job.w, job.g, a.workload, threads, sample)
            r1 = cpu_reference_run(job.w, job.g, a.workload, 1, max(2, sample // (4 * threads)))
            cpu = dict(value=rc["out_bytes"] / rc["secs"] / 1e9, unit="GB/s", cores=threads, kind="reference",
                       one_core_value=r1["out_bytes"] / r1["secs"] / 1e9, sample_units=rc["units"],
                       sample=f"{rc['units']} units ({rc['out_bytes'] / 2**20:.0f} MiB out), one unit per thread, unmodified "
                              f"reference decoder (oracle/_ref/libref.so), probs allocated once per thread")
            log(f"cpu baseline {cpu['value']:.3f} GB/s on {threads} threads")
    ratio = job.g["sum_p"] / job.g["sum_u"]
    del job
    lzb200.sdk.pinned_free_all()
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configs in the same run
    if a.extras == "auto":
        extras = ["weak_lzma64k", "xz4m", "7z", "lzma_stream"] if world == 1 else ["weak_lzma64k"]
        if a.small or a.workload != "lzma2_8g":
            extras = []
    elif a.extras in ("none", ""):
        extras = []
    elif a.extras == "all":
        extras = ["weak_lzma64k", "xz4m", "7z", "lzma_stream"]
    else:
        extras = [x for x in a.extras.split(",") if x]
    configs = {}
    for name in extras:
        t0 = time.time()
        try:
            if name in ("weak_lzma64k", "xz4m"):
                wl = "lzma64k" if name == "weak_lzma64k" else "xz4m"
                j = Job(wl, DEFAULT_UNITS[wl] // (16 if a.small else 1), a, env, replicate=(name == "weak_lzma64k"))
                j.stage()
                rr = j.run(2, 3, e2e=not a.no_e2e)
                d = dict(workload=WORKLOADS[wl], units_per_gpu=rr["units_this_rank"], unique_units=len(j.w["plen"]),
                         scaling="weak (replicated per GPU)" if name == "weak_lzma64k" else "strong",
                         value=rr["value"], unit="GB/s", ms_per_step=rr["ms_per_step"], verified=rr["verified"], e2e=rr["e2e"],
                         roofline=roofline_of(wl, rr, geometry_note(rr["units_this_rank"], ctx.sms or 148)))
                if rank == 0 and world == 1 and not a.no_cpu:
                    s = cpu_sample_units(wl, threads, len(j.g["units"]))
                    rc = cpu_reference_run(j.w, j.g, wl, threads, s)
                    d["cpu_baseline"] = dict(value=rc["out_bytes"] / rc["secs"] / 1e9, unit="GB/s", cores=threads, kind="reference",
                                             sample_units=rc["units"], sample=f"{rc['units']} units, one per thread")
                del j
                lzb200.sdk.pinned_free_all()
                torch.cuda.empty_cache()
            elif rank == 0:
                d = extra_7z(env) if name == "7z" else extra_lzma_stream(env)
            else:
                d = None
            if d is not None:
                d["bench_seconds"] = round(time.time() - t0, 1)
                configs[name] = d
                log(f"extra {name}: {d.get('value')} GB/s in {d['bench_seconds']} s")
        except Exception as e:                                   # an extra must never take the headline line down
            configs[name] = dict(error=f"{type(e).__name__}: {e}")
            log(f"extra {name} failed: {e}")
        if dist:
            dist.barrier()

    if rank == 0:
        line = dict(metric=METRIC, value=r["value"], unit="GB/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak" if weak else "strong", vs_baseline=None,
                    dtype="u8", data="synthetic", config=cfg, clocks=r["clocks"], e2e=r["e2e"], gpu_launches=a.steps * world,
                    roofline=roofline, cpu_baseline=cpu, verified=r["verified"],
                    ratio=dict(packed_over_unpacked=ratio))
        if container:
            line["container_e2e"] = container
        if r.get("gather"):
            line["gather"] = r["gather"]
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
