#!/usr/bin/env python
"""bench.py -- decompressed GB/s of the B200 decode path next to the reference's CPU decode.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lzma64k|lzma2_1m|xz4m] [--impl reference]

Workload (default = BASELINE.json configs[1]): a batch of 65,536 independent 64 KiB raw LZMA streams
(LzmaUncompress contract, level 5, lc3 lp0 pb2, 64 KiB dictionary), synthetic mixed-entropy corpus
compressed by the reference encoder.  To keep generation within seconds the batch is tiled from 2,048
uniquely encoded streams; every tile is a physical copy (packed arena ~2.9 GB, output 4.3 GB per GPU --
far larger than the 126 MB L2, so no step is served from cache).  A step is one pass over the batch.

  value  : device-resident throughput -- arenas already in HBM, one fused decode launch per step
           (Lzb_DecodeUnitsDevice), timed with CUDA events on the launching stream.
  e2e    : the same batch through the C-ABI with HOST (pinned) arenas, H2D and D2H inside the timed
           region (Lzb_DecodeUnitsHost).
  N > 1  : one process per GPU (torchrun), units sharded statically, no data-path collective;
           per-GPU work is fixed ("weak"); time = max over ranks, value = all ranks' bytes / that time.
  --impl reference : the unmodified reference decoder (oracle/_ref/libref.so, one unit per host thread)
           on a bounded sample of the same batch.
Extra keys beside the contract's: `issue_roofline` (the bound that actually limits the kernel: issued warp
instructions per SM and cycle, of 4), `e2e.per_call_ms`, and for --workload xz4m `container_e2e` (the same blocks
as ONE .xz file through XzDecodeParallel).
"""
from __future__ import annotations

import argparse
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

from tools.corpus import add_x86_callsites, make_corpus  # noqa: E402

METRIC = "decompressed GB/s"
WORKLOADS = {
    "lzma64k": "65,536 independent 64 KiB raw LZMA streams (LzmaUncompress), level 5, lc3 lp0 pb2, dict 64 KiB",
    "lzma2_1m": "LZMA2 stream of independent 1 MiB units (dict reset each), level 1, dict 1 MiB, Lzma2Dec contract",
    "xz4m": "xz container, 4 MiB blocks, x86 BCJ + LZMA2 + CRC64 per block (XzDec block path)",
}


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


# ------------------------------------------------------------------------------------ workload
def build_workload(name: str, rank: int, cache_dir: str, small: bool):
    """Returns dict(packed arena, unit table fields, expected per-unit CRC32, props...).  Rank 0 encodes
    (reference encoder, all host threads) and caches under cache_dir; other ranks load the cache."""
    from oracle import refbind
    os.makedirs(cache_dir, exist_ok=True)
    path = os.path.join(cache_dir, f"{name}{'_small' if small else ''}.npz")
    if not os.path.exists(path):
        if rank != 0:
            raise RuntimeError("cache missing on non-zero rank")
        ref = refbind.ref()
        thr = host_threads()
        if name == "lzma64k":
            uniq, unit = (256, 65536) if small else (2048, 65536)
            corpus = make_corpus(uniq * unit, seed=1234)
            offs = np.arange(uniq, dtype=np.uint64) * unit
            lens = np.full(uniq, unit, dtype=np.uint32)
            enc = ref.encode_batch_mt(0, corpus, offs, lens, level=5, dict_size=1 << 16, lc=3, lp=0, pb=2, fb=32,
                                      end_mark=0, threads=thr)
            kind, props = 0, enc["props"]
        elif name == "lzma2_1m":
            uniq, unit = (16, 1 << 20) if small else (128, 1 << 20)
            corpus = make_corpus(uniq * unit, seed=1234)
            offs = np.arange(uniq, dtype=np.uint64) * unit
            lens = np.full(uniq, unit, dtype=np.uint32)
            enc = ref.encode_batch_mt(1, corpus, offs, lens, level=1, dict_size=1 << 20, lc=3, lp=0, pb=2, threads=thr)
            kind, props = 1, enc["props"]
        else:
            uniq, unit = (4, 4 << 20) if small else (32, 4 << 20)
            corpus = add_x86_callsites(make_corpus(uniq * unit, seed=1234), seed=7)
            filt = np.empty_like(corpus)
            for k in range(uniq):
                filt[k * unit:(k + 1) * unit] = ref.x86_filter_whole(corpus[k * unit:(k + 1) * unit], 0, 1)
            offs = np.arange(uniq, dtype=np.uint64) * unit
            lens = np.full(uniq, unit, dtype=np.uint32)
            enc = ref.encode_batch_mt(1, filt, offs, lens, level=1, dict_size=4 << 20, lc=3, lp=0, pb=2, threads=thr)
            kind, props = 2, enc["props"]
        crcs = np.array([zlib.crc32(corpus[int(o):int(o) + int(l)].tobytes()) for o, l in zip(offs, lens)], dtype=np.uint64)
        crc64 = np.zeros(uniq, dtype=np.uint64)
        if kind == 2:
            crc64 = np.array([ref.crc64(corpus[int(o):int(o) + int(l)]) for o, l in zip(offs, lens)], dtype=np.uint64)
        blob = np.concatenate([enc["buf"][int(o):int(o) + int(l)] for o, l in zip(enc["off"], enc["len"])])
        np.savez(path + ".tmp.npz", blob=blob, plen=enc["len"], ulen=lens, crc32=crcs, crc64=crc64,
                 props=np.asarray(props, dtype=np.uint8), kind=kind, encode_secs=enc["secs"])
        os.replace(path + ".tmp.npz", path)
    z = np.load(path)
    return {k: z[k] for k in z.files}


def tile_units(w, n_units: int, name: str):
    """Physical tiling of the unique streams into the full batch: arenas + unit table (numpy)."""
    import lzb200
    R = lzb200.records
    uniq = len(w["plen"])
    idx = np.arange(n_units) % uniq
    plen = w["plen"][idx].astype(np.uint64)
    ulen = w["ulen"][idx].astype(np.uint64)
    pad = (plen + 15) & ~np.uint64(15)
    soff = np.concatenate(([0], np.cumsum(pad)[:-1])).astype(np.uint64)
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
    total_p = int(soff[-1] + pad[-1]) + 64
    total_u = int(doff[-1] + ulen[-1])
    return dict(idx=idx, units=units, soff=soff, plen=plen, doff=doff, ulen=ulen, ustart=ustart,
                packed_bytes=total_p, out_bytes=total_u, sum_p=int(plen.sum()), sum_u=int(ulen.sum()))


def fill_packed(dst: np.ndarray, w, t):
    blob = w["blob"]
    for i in range(len(t["idx"])):
        k = int(t["idx"][i])
        s, n = int(t["ustart"][k]), int(w["plen"][k])
        o = int(t["soff"][i])
        dst[o:o + n] = blob[s:s + n]


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
def cpu_reference_run(w, t, name: str, threads: int, sample_units: int, repeat: int = 1):
    """The unmodified reference decoder, one unit per host thread pulled from a shared counter."""
    from oracle import refbind
    ref = refbind.ref()
    n = min(sample_units, len(t["idx"]))
    kind = int(w["kind"])
    best = None
    packed = np.zeros(int(t["soff"][n - 1] + t["plen"][n - 1]) + 64, dtype=np.uint8)
    sub = dict(t)
    sub["idx"] = t["idx"][:n]
    fill_packed(packed, w, sub)
    if kind == 2:
        # xz leg: each block wrapped as a one-block .xz member so the whole XzUnpacker path (block header,
        # LZMA2, BCJ, CRC-64, index, footer) is timed, one member per host thread
        uniq = len(w["plen"])
        ustart = np.concatenate(([0], np.cumsum(w["plen"].astype(np.uint64))[:-1]))
        members = [ref.xz_assemble(w["blob"][int(ustart[k]):int(ustart[k]) + int(w["plen"][k])], [0], [int(w["plen"][k])],
                                   [int(w["ulen"][k])], [int(w["crc64"][k])], 4, 4, 0, int(w["props"][0])) for k in range(uniq)]
        idx = t["idx"][:n]
        mlen = np.array([len(members[int(k)]) for k in idx], dtype=np.uint32)
        moff = np.concatenate(([0], np.cumsum((mlen.astype(np.uint64) + 15) & ~np.uint64(15))[:-1])).astype(np.uint64)
        arena = np.zeros(int(moff[-1]) + int(mlen[-1]) + 64, dtype=np.uint8)
        for i, k in enumerate(idx):
            arena[int(moff[i]):int(moff[i]) + int(mlen[i])] = members[int(k)]
        for _ in range(repeat):
            r = ref.xz_decode_batch_mt(arena, moff, mlen, t["doff"][:n], t["ulen"][:n].astype(np.uint32), threads, scratch=True)
            assert (r["res"] == 0).all() and (r["dst_done"] == t["ulen"][:n]).all(), "reference xz decode failed"
            best = r["secs"] if best is None else min(best, r["secs"])
        return dict(secs=best, out_bytes=int(t["ulen"][:n].sum()), units=n)
    for _ in range(repeat):
        r = ref.decode_batch_mt(0 if kind == 0 else 1, packed, t["soff"][:n], t["plen"][:n].astype(np.uint32),
                                t["doff"][:n], t["ulen"][:n].astype(np.uint32), w["props"],
                                0 if kind == 0 else 1, threads, scratch=True)
        assert (r["res"] == 0).all() and (r["dst_done"] == t["ulen"][:n]).all(), "reference decode failed"
        secs = r["secs"]
        best = secs if best is None else min(best, secs)
    out_bytes = int(t["ulen"][:n].sum())
    return dict(secs=best, out_bytes=out_bytes, units=n)


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lzma64k", choices=sorted(WORKLOADS))
    ap.add_argument("--units", type=int, default=0, help="units per GPU (default: the BASELINE size of the workload)")
    ap.add_argument("--small", action="store_true", help="tiny corpus, for smoke-testing the harness")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    cache = os.environ.get("LZB_BENCH_CACHE", "/tmp/lzb_bench_cache")
    default_units = {"lzma64k": 65536, "lzma2_1m": 8192, "xz4m": 2220}[a.workload]   # xz4m: one 4 MiB block per resident warp
    n_units = a.units or (default_units // 16 if a.small else default_units)
    threads = host_threads()
    cfg = dict(workload=WORKLOADS[a.workload], units_per_gpu=n_units, sharding="static, units split by rank, no collective",
               l2="inputs >> L2 (126 MB): every tile is a physical copy", corpus="synthetic mixed-entropy, seed 1234, "
               "reference encoder; batch tiled from uniquely encoded streams")

    # ---------------- reference arm: CPU only, rank 0 only
    if a.impl == "reference":
        if rank != 0:
            return 0
        w = build_workload(a.workload, 0, cache, a.small)
        t = tile_units(w, n_units, a.workload)
        sample = min(n_units, max(64, (512 if a.workload == "lzma64k" else 8) * threads))
        for _ in range(max(a.warmup, 0) and 1):
            cpu_reference_run(w, t, a.workload, threads, min(sample, 4 * threads))
        times = [cpu_reference_run(w, t, a.workload, threads, sample) for _ in range(a.steps)]
        secs = float(np.mean([x["secs"] for x in times]))
        v = times[0]["out_bytes"] / secs / 1e9
        line = dict(metric=METRIC, value=v, unit="GB/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=secs * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8",
                    data="synthetic", config=cfg, impl="reference",
                    cpu_baseline=dict(value=v, unit="GB/s", cores=threads, kind="reference",
                                      sample=f"{times[0]['units']} units ({times[0]['out_bytes'] / 2**20:.0f} MiB out) per step, "
                                             f"one unit per thread, unmodified LzmaDec/Lzma2Dec/XzUnpacker (oracle/_ref/libref.so)"),
                    e2e=dict(value=v, unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    # ---------------- B200 arm
    import torch
    import lzb200
    from lzb200 import records as R
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        w = build_workload(a.workload, 0, cache, a.small)
    if dist:
        dist.barrier()
    if rank != 0:
        w = build_workload(a.workload, rank, cache, a.small)
    t = tile_units(w, n_units, a.workload)
    units = t["units"]
    max_lclp, max_pb = int((units["lc"].astype(int) + units["lp"]).max()), int(units["pb"].max())

    ctx = lzb200.lib.Context(local)
    h_packed = lzb200.sdk.pinned_empty(t["packed_bytes"])
    fill_packed(h_packed, w, t)
    h_out = lzb200.sdk.pinned_empty(t["out_bytes"] + 64)
    dev = torch.device("cuda", local)
    d_packed = torch.empty(t["packed_bytes"], dtype=torch.uint8, device=dev)
    d_packed.copy_(torch.from_numpy(h_packed), non_blocking=False)
    d_out = torch.empty(t["out_bytes"] + 64, dtype=torch.uint8, device=dev)
    d_units = torch.from_numpy(units.view(np.uint8).reshape(-1)).to(dev)
    d_results = torch.zeros(n_units * R.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)      # the stream the decode kernels are launched on AND timed on
    torch.cuda.synchronize()

    def step_device():
        lzb200.sdk.decode_units_device(ctx, d_packed.data_ptr(), d_out.data_ptr(), d_units.data_ptr(),
                                       d_results.data_ptr(), n_units, max_lclp, max_pb, stream.cuda_stream)

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

    for _ in range(a.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record(stream)
    for k in range(a.steps):
        step_device()
        ev[k + 1].record(stream)
    sync_all()
    clocks = sampler.stop()
    total_ms = reduce_max(ev[0].elapsed_time(ev[-1]))
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)]
    ms_per_step = total_ms / a.steps
    out_bytes_all = t["sum_u"] * world
    value = out_bytes_all / (ms_per_step * 1e-3) / 1e9

    # correctness of what was just timed (outside the timed region): statuses + digest of digests
    res = d_results.cpu().numpy().view(R.RESULT_DTYPE)
    ok = bool((res["res"] == 0).all() and (res["dst_done"] == t["ulen"]).all())
    vunits = units.copy()
    if int(w["kind"]) != 2:
        vunits["check"], vunits["flags"], vunits["expect"] = R.CHECK_CRC32, R.UNIT_VERIFY, w["crc32"][t["idx"]]
    d_vunits = torch.from_numpy(vunits.view(np.uint8).reshape(-1)).to(dev)
    lzb200.sdk.decode_units_device(ctx, d_packed.data_ptr(), d_out.data_ptr(), d_vunits.data_ptr(), d_results.data_ptr(),
                                   n_units, max_lclp, max_pb, stream.cuda_stream)
    torch.cuda.synchronize()
    vres = d_results.cpu().numpy().view(R.RESULT_DTYPE)
    verified = ok and bool((vres["res"] == 0).all())
    del d_vunits

    # ---------------- e2e: host arenas through the C-ABI, copies inside the timed region
    e2e = None
    if not a.no_e2e:
        for _ in range(max(a.warmup, 1)):
            lzb200.sdk.decode_units_host(ctx, h_packed, h_out, units)
        sync_all()
        t0 = time.perf_counter()
        per_call = []
        for _ in range(a.steps):
            tc = time.perf_counter()
            hres = lzb200.sdk.decode_units_host(ctx, h_packed, h_out, units)
            per_call.append((time.perf_counter() - tc) * 1e3)
        torch.cuda.synchronize()

        dt = reduce_max(time.perf_counter() - t0)
        if dist:
            dist.barrier()
        e2e_ok = bool((hres["res"] == 0).all())
        k = n_units // 2
        e2e_ok = e2e_ok and zlib.crc32(h_out[int(t["doff"][k]):int(t["doff"][k] + t["ulen"][k])].tobytes()) == int(w["crc32"][t["idx"][k]])
        e2e = dict(value=out_bytes_all / (dt / a.steps) / 1e9, unit="GB/s",
                   h2d_bytes_per_step=int(t["sum_p"] + units.nbytes), d2h_bytes_per_step=int(t["sum_u"] + res.nbytes),
                   ms_per_step=dt / a.steps * 1e3, per_call_ms=[round(x, 2) for x in per_call], verified=e2e_ok,
                   api="Lzb_DecodeUnitsHost (pinned host arenas; one launch fed and drained while it runs)")

    # ---------------- xz4m only: the same blocks as ONE real .xz file through the container API (rank 0)
    container = None
    if a.workload == "xz4m" and not a.no_e2e and rank == 0 and world == 1:
        from oracle import refbind
        if refbind.have_ref():
            ref = refbind.ref()
            nb = n_units
            idx = t["idx"][:nb]
            ustart = np.concatenate(([0], np.cumsum(w["plen"].astype(np.uint64))[:-1]))
            plen = w["plen"][idx].astype(np.uint32)
            poff = np.concatenate(([0], np.cumsum(plen.astype(np.uint64))[:-1])).astype(np.uint64)
            blob = np.concatenate([w["blob"][int(ustart[k]):int(ustart[k]) + int(w["plen"][k])] for k in idx])
            xz = ref.xz_assemble(blob, poff, plen, w["ulen"][idx].astype(np.uint32), w["crc64"][idx], 4, 4, 0, int(w["props"][0]))
            total = int(w["ulen"][idx].astype(np.uint64).sum())
            hx = lzb200.sdk.pinned_empty(len(xz)); hx[:] = xz
            ho = lzb200.sdk.pinned_empty(total + 16)
            times = []
            for i in range(2 + a.steps):
                tc = time.perf_counter()
                d = lzb200.sdk.xz_decode(hx, total, ctx, out=ho)
                if i >= 2:
                    times.append(time.perf_counter() - tc)
            okc = d["res"] == 0 and d["finished"] and d["dest_len"] == total and d["num_blocks"] == nb and \
                zlib.crc32(d["out"][:int(w["ulen"][idx[0]])].tobytes()) == int(w["crc32"][idx[0]])
            # the reference on one thread: timed on the first 32 blocks as a file of their own (bounded sample)
            ns = min(nb, 32)
            xs = ref.xz_assemble(blob[: int(poff[ns - 1]) + int(plen[ns - 1])], poff[:ns], plen[:ns], w["ulen"][idx[:ns]].astype(np.uint32),
                                 w["crc64"][idx[:ns]], 4, 4, 0, int(w["props"][0]))
            tot_s = int(w["ulen"][idx[:ns]].astype(np.uint64).sum())
            t1 = time.perf_counter()
            rr = ref.xz_decode(xs, tot_s)
            cpu1 = time.perf_counter() - t1
            container = dict(api="XzDecodeParallel on one .xz file (host buffers; framing on the host, all blocks in one GPU pass)",
                             blocks=nb, file_bytes=int(len(xz)), out_bytes=total, value=total / float(np.mean(times)) / 1e9, unit="GB/s",
                             verified=bool(okc), reference_one_thread=dict(value=tot_s / cpu1 / 1e9, unit="GB/s", res=int(rr["res"]), blocks=ns,
                             note="the reference decodes one .xz file on one thread (XzUnpacker_Code loop)"))

    # ---------------- roofline of the dominant kernel (the fused decode launch = the whole step)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kms = float(np.mean(kernel_ms))
    achieved = (t["sum_p"] + t["sum_u"]) / (kms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            rec = json.load(open(tp)).get(a.workload)
            if rec:   # measured DRAM bytes per algorithmic byte (ncu --set full), scaled to this launch
                traffic = int(rec["dram_bytes_per_algorithmic_byte"] * (t["sum_p"] + t["sum_u"]))
        except Exception:
            traffic = None
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                    kernel="lzb_decode_units_kernel<shared model, G lanes per unit>", launches_per_step=1, kernel_ms=kms,
                    algorithmic_bytes_per_launch=int(t["sum_p"] + t["sum_u"]), peak_source=peak_src,
                    note="serial range-decoder chain: bound by resident units x issue slots, far below the HBM roof; "
                         "traffic above the algorithmic bytes = match sources evicted from L2 (profiles/r01_c_residency.md)")

    # what actually bounds the kernel: issue slots.  Warp instructions per output byte is a
    # property of the code + corpus measured by ncu (profiles/r01_f_final.md); the rest is measured live here.
    issue = None
    if a.workload == "lzma64k" and clocks.get("sm_mhz"):
        wi_per_byte = 107.6
        ipc = wi_per_byte * t["sum_u"] / (kms * 1e-3) / (clocks["sm_mhz"] * 1e6) / float(ctx.sms or 148)
        issue = dict(bound="issue slots", warp_instr_per_output_byte=wi_per_byte, achieved_ipc_per_sm=ipc, peak_ipc_per_sm=4.0, frac=ipc / 4.0,
                     source="ncu smsp__inst_executed.sum / output bytes (profiles/r01_f_final.md) x live kernel time and SM clock")

    # ---------------- CPU baseline on this box (rank 0, N=1 only), bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import refbind
        if refbind.have_ref():
            wc = dict(w)
            sample = min(n_units, max(64, (512 if a.workload == "lzma64k" else 8) * threads))
            cpu_reference_run(wc, t, a.workload, threads, min(sample, 4 * threads))
            r = cpu_reference_run(wc, t, a.workload, threads, sample)
            r1 = cpu_reference_run(wc, t, a.workload, 1, max(4, sample // (4 * threads)))
            cpu = dict(value=r["out_bytes"] / r["secs"] / 1e9, unit="GB/s", cores=threads, kind="reference",
                       one_core_value=r1["out_bytes"] / r1["secs"] / 1e9,
                       sample=f"{r['units']} units ({r['out_bytes'] / 2**20:.0f} MiB out), one unit per thread, unmodified "
                              f"reference decoder (oracle/_ref/libref.so), probs allocated once per thread")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit="GB/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8",
                    data="synthetic", config=cfg, clocks=clocks, e2e=e2e, gpu_launches=a.steps, roofline=roofline,
                    cpu_baseline=cpu, verified=verified,
                    ratio=dict(packed_over_unpacked=t["sum_p"] / t["sum_u"]))
        if container:
            line["container_e2e"] = container
        if issue:
            line["issue_roofline"] = issue
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
