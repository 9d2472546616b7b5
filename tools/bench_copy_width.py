#!/usr/bin/env python
"""Copy width where it matters (north_star: "the copy width must be justified"): units whose output is almost all
long byte runs (the (0,0,1) end of config 5's entropy sweep: rep0 = 1 matches of up to 273 bytes) and units that are
stored LZMA2 chunks (random data: the window copy is a plain memcpy).  Device-resident, CUDA events.

    python tools/bench_copy_width.py                      (default library: 16 bytes per lane for runs)
    LZB_LIB=.../variants/copybytes/liblzb200.so python tools/bench_copy_width.py     (one byte per lane)
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.corpus import make_corpus  # noqa: E402


def main():
    import torch
    import lzb200
    from oracle import refbind
    R = lzb200.records
    ref = refbind.ref()
    ctx = lzb200.lib.Context(0)
    dev = torch.device("cuda", 0)
    thr = max(1, len(os.sched_getaffinity(0)))
    out = dict(tool="tools/bench_copy_width.py", lib=os.environ.get("LZB_LIB", "default"))
    for name, mix, kind in (("byte_runs", (0.0, 0.0, 1.0), 0), ("random_stored_lzma2", (1.0, 0.0, 0.0), 1), ("mixed", (0.25, 0.5, 0.25), 0)):
        uniq, unit, n_units = 512, 65536, 32768
        corpus = make_corpus(uniq * unit, seed=77, mix=mix)
        offs = np.arange(uniq, dtype=np.uint64) * unit
        lens = np.full(uniq, unit, dtype=np.uint32)
        enc = ref.encode_batch_mt(kind, corpus, offs, lens, level=1, dict_size=1 << 16, lc=3, lp=0, pb=2, fb=32, end_mark=0, threads=thr)
        idx = np.arange(n_units) % uniq
        plen = enc["len"][idx].astype(np.uint64)
        pad = (plen + 15) & ~np.uint64(15)
        soff = np.concatenate(([0], np.cumsum(pad)[:-1])).astype(np.uint64)
        packed = np.zeros(int(soff[-1] + pad[-1]) + 64, np.uint8)
        for i in range(n_units):
            k = int(idx[i])
            packed[int(soff[i]):int(soff[i]) + int(enc["len"][k])] = enc["buf"][int(enc["off"][k]):int(enc["off"][k]) + int(enc["len"][k])]
        u = R.make_units(n_units)
        u["src_off"], u["src_len"], u["dst_off"], u["dst_cap"] = soff, plen, np.arange(n_units, dtype=np.uint64) * unit, unit
        if kind == 0:
            lc, lp, pb, ds = R.parse_lzma_props(enc["props"])
            u["dict_size"], u["codec"], u["lc"], u["lp"], u["pb"], u["finish"] = ds, R.CODEC_LZMA, lc, lp, pb, 0
        else:
            prop = int(enc["props"][0])
            u["dict_size"], u["codec"], u["lc"], u["pb"], u["finish"] = (2 | (prop & 1)) << (prop // 2 + 11), R.CODEC_LZMA2, 3, 2, 1
        u["check"], u["flags"] = R.CHECK_CRC32, R.UNIT_VERIFY
        u["expect"] = np.array([zlib.crc32(corpus[int(o):int(o) + unit].tobytes()) for o in offs], dtype=np.uint64)[idx]
        d_p = torch.from_numpy(packed).to(dev)
        d_o = torch.empty(n_units * unit + 64, dtype=torch.uint8, device=dev)
        d_u = torch.from_numpy(u.view(np.uint8).reshape(-1)).to(dev)
        d_r = torch.zeros(n_units * 24, dtype=torch.uint8, device=dev)
        st = torch.cuda.Stream(device=dev)
        lclp, pbm = int((u["lc"].astype(int) + u["lp"]).max()), int(u["pb"].max())
        for _ in range(3):
            lzb200.sdk.decode_units_device(ctx, d_p.data_ptr(), d_o.data_ptr(), d_u.data_ptr(), d_r.data_ptr(), n_units, lclp, pbm, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            lzb200.sdk.decode_units_device(ctx, d_p.data_ptr(), d_o.data_ptr(), d_u.data_ptr(), d_r.data_ptr(), n_units, lclp, pbm, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res = d_r.cpu().numpy().view(R.RESULT_DTYPE)
        out[name] = dict(out_GBps=n_units * unit / ms / 1e6, ms=ms, packed_over_unpacked=float(plen.sum()) / (n_units * unit),
                         verified=bool((res["res"] == 0).all()), out_plus_in_GBps=(n_units * unit + float(plen.sum())) / ms / 1e6)
        del d_p, d_o, d_u, d_r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
