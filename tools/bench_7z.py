"""BASELINE config 5 through the container API: a 7z archive of many small files in many folders, from low
entropy to random -- SzArchive_Open + SzArchive_ExtractAll (liblzb200.so, host buffers, every folder in one GPU
pass, folder and file CRC-32 on the device) next to the reference's SzArEx_Open + SzArEx_Extract loop on one
host thread (how the reference extracts an archive).  Prints one JSON line; not the driver's bench contract
(that is bench.py) -- the line is kept under profiles/.

    python tools/bench_7z.py [--folders 8192] [--file-kib 32] [--steps 3]
"""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools import sz_writer as W  # noqa: E402
from tools.corpus import entropy_sweep_mix, make_corpus  # noqa: E402


def build(ref, n_folders, file_bytes, threads):
    """n_folders single-file folders; file i is drawn from a corpus whose mix moves from runs to random."""
    parts = []
    for k in range(16):
        parts.append(make_corpus(n_folders // 16 * file_bytes, seed=500 + k, mix=entropy_sweep_mix(k / 15.0)))
    corpus = np.concatenate(parts)
    n = len(corpus) // file_bytes
    offs = np.arange(n, dtype=np.uint64) * file_bytes
    lens = np.full(n, file_bytes, dtype=np.uint32)
    enc = ref.encode_batch_mt(0, corpus, offs, lens, level=1, dict_size=1 << 16, lc=3, lp=0, pb=2, fb=32, end_mark=0, threads=threads)
    folders = []
    for i in range(n):
        data = corpus[i * file_bytes:(i + 1) * file_bytes].tobytes()
        f = W.Folder("lzma", [(f"d{i % 97:02d}/f{i:06d}.bin", data)])
        f.packed = enc["buf"][int(enc["off"][i]):int(enc["off"][i]) + int(enc["len"][i])].tobytes()
        f.props = enc["props"].tobytes()
        f.plain = data
        folders.append(f)
    return W.write_archive(ref, folders, encode_header=True), corpus


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--folders", type=int, default=8192)
    ap.add_argument("--file-kib", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    from oracle import refbind
    import lzb200
    ref = refbind.ref()
    threads = max(1, len(os.sched_getaffinity(0)))
    arc, corpus = build(ref, a.folders, a.file_kib << 10, threads)
    arc = np.frombuffer(arc, dtype=np.uint8)
    d = lzb200.sdk.sz_extract_all(arc)          # full result (file table, names, bytes) for the check below
    # timed: the C calls a C host would make -- open (database + packed header on the GPU), extract all, free
    import ctypes as C
    L = lzb200.lib.load()
    n_files = len(d["files"])
    out = lzb200.sdk.pinned_empty(len(corpus) + 64)
    harc = lzb200.sdk.pinned_empty(len(arc)); harc[:] = arc
    fres, ires = np.zeros(a.folders + 8, np.int32), np.zeros(n_files + 8, np.int32)
    times = []
    for i in range(2 + a.steps):
        t0 = time.perf_counter()
        h = C.c_void_p(None)
        r1 = L.SzArchive_Open(None, harc.ctypes.data, len(harc), C.byref(h))
        r2 = L.SzArchive_ExtractAll(None, h, harc.ctypes.data, len(harc), out.ctypes.data, len(corpus), fres.ctypes.data, ires.ctypes.data)
        L.SzArchive_Free(h)
        if i >= 2:
            times.append(time.perf_counter() - t0)
        assert r1 == 0 and r2 == 0 and not fres[: a.folders].any() and not ires[:n_files].any()
    assert zlib.crc32(out[: len(corpus)].tobytes()) == zlib.crc32(corpus.tobytes())
    ok = d["open_res"] == 0 and all(f["res"] == 0 for f in d["files"]) and len(d["files"]) == len(corpus) // (a.file_kib << 10) and \
        zlib.crc32(d["out"].tobytes()) == zlib.crc32(corpus.tobytes())
    total = int(len(corpus))
    # the reference: one thread, bounded sample (an archive of the first 512 folders)
    sub_n = min(512, a.folders)
    sub, _ = build(ref, sub_n, a.file_kib << 10, threads) if sub_n != a.folders else (arc.tobytes(), None)
    t0 = time.perf_counter()
    r = ref.sz_extract_all(sub, out_cap=sub_n * (a.file_kib << 10) + 4096, max_files=sub_n + 8)
    cpu = time.perf_counter() - t0
    assert r["open_res"] == 0 and all(f["res"] == 0 for f in r["files"])
    line = dict(metric="decompressed GB/s", workload=f"7z archive, {a.folders} single-file LZMA folders of {a.file_kib} KiB, entropy sweep, "
                "packed header, folder CRC-32 (BASELINE config 5)", api="SzArchive_Open + SzArchive_ExtractAll + SzArchive_Free (pinned host buffers)",
                value=total / float(np.mean(times)) / 1e9, unit="GB/s", ms_per_call=[round(x * 1e3, 1) for x in times], archive_bytes=int(len(arc)),
                out_bytes=total, verified=bool(ok),
                reference_one_thread=dict(value=sub_n * (a.file_kib << 10) / cpu / 1e9, unit="GB/s", folders=sub_n,
                                          api="SzArEx_Open + SzArEx_Extract per file (oracle/_ref/libref.so)"))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
