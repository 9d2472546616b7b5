#!/usr/bin/env python
"""What a caller that does NOT batch gets from the stock one-call API ("unbatched callers"):

  * latency of ONE LzmaUncompress of a 64 KiB stream, CrcCalc of 4 KiB .. 4 MiB, x86_Convert of 1 MiB -- every call is
    host->device, one launch on one warp (or a few), device->host -- next to the reference on one host core;
  * the same LzmaUncompress called from 1 .. 64 host threads at once: concurrent calls are combined into one launch
    (lzb_host.cu, "callers that do not batch"), so the rate grows with the number of callers.

    python tools/bench_unbatched.py  > profiles/r02_unbatched.json        (needs a GPU)
"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.corpus import add_x86_callsites, make_corpus  # noqa: E402


def timed(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return float(np.median(t))


def main():
    import lzb200
    from oracle import refbind
    ref = refbind.ref()
    L = lzb200.lib.load()
    S = lzb200.sdk
    out = dict(tool="tools/bench_unbatched.py", note="medians; GPU calls include both copies and the launch")
    corpus = make_corpus(64 * 65536, seed=99)
    streams = []
    for i in range(64):
        pk, props = ref.lzma_encode(corpus[i * 65536:(i + 1) * 65536], level=5, dict_size=1 << 16)
        streams.append((np.ascontiguousarray(pk), np.ascontiguousarray(props)))
    pk, props = streams[0]
    S.lzma_uncompress(pk, props, 65536)                       # context creation, first launch
    g = timed(lambda: S.lzma_uncompress(pk, props, 65536), 20)
    c = timed(lambda: ref.lzma_decode(pk, props, 65536, 0), 20)
    out["LzmaUncompress_64KiB"] = dict(gpu_ms=g * 1e3, reference_one_core_ms=c * 1e3, gpu_MBps=65536 / g / 1e6, reference_MBps=65536 / c / 1e6)
    crc = {}
    for n in (4 << 10, 64 << 10, 1 << 20, 4 << 20):
        buf = corpus[:n] if n <= len(corpus) else np.resize(corpus, n)
        g = timed(lambda: S.crc32(buf), 20)
        c = timed(lambda: ref.crc32(buf), 20)
        crc[str(n)] = dict(gpu_ms=g * 1e3, reference_one_core_ms=c * 1e3)
    out["CrcCalc"] = crc
    x = add_x86_callsites(make_corpus(1 << 20, seed=5), seed=6)
    g = timed(lambda: S.x86_convert(x, 0, 0, 0), 20)
    c = timed(lambda: ref.x86_convert(x, 0, 0, 0), 20)
    out["x86_Convert_1MiB"] = dict(gpu_ms=g * 1e3, reference_one_core_ms=c * 1e3)

    # resumable decoder object: cost of ONE LzmaDec_DecodeToBuf call that has next to nothing to decode (16 bytes out):
    # two small uploads, one launch, three small downloads, one wait (lzb_state.cu)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from streamdrive import CLzmaDec, _alloc
    L.LzmaDec_Allocate.restype = C.c_int
    L.LzmaDec_Allocate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
    L.LzmaDec_Init.argtypes = [C.c_void_p]
    L.LzmaDec_Free.argtypes = [C.c_void_p, C.c_void_p]
    L.LzmaDec_DecodeToBuf.restype = C.c_int
    L.LzmaDec_DecodeToBuf.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_int)]
    d = CLzmaDec()
    assert L.LzmaDec_Allocate(C.byref(d), props.ctypes.data, 5, C.byref(_alloc)) == 0
    L.LzmaDec_Init(C.byref(d))
    dst = np.zeros(65536, np.uint8)
    ip = op = 0
    st = C.c_int(0)
    per_call = []
    ok_state = True
    while op < 65536 and len(per_call) < 400:
        dl, sl = C.c_size_t(16), C.c_size_t(min(64, len(pk) - ip))
        t0 = time.perf_counter()
        r = L.LzmaDec_DecodeToBuf(C.byref(d), dst[op:].ctypes.data, C.byref(dl), pk[ip:].ctypes.data, C.byref(sl), 0, C.byref(st))
        per_call.append(time.perf_counter() - t0)
        ip += sl.value
        op += dl.value
        if r != 0:
            ok_state = False
            break
    ok_state = ok_state and np.array_equal(dst[:op], corpus[:op])
    L.LzmaDec_Free(C.byref(d), C.byref(_alloc))
    out["LzmaDec_DecodeToBuf_16_bytes_per_call"] = dict(calls=len(per_call), median_us=float(np.median(per_call[5:])) * 1e6,
                                                        p90_us=float(np.percentile(per_call[5:], 90)) * 1e6, ok=bool(ok_state))

    # N host threads, each calling LzmaUncompress in a loop on its own stream for ~2 s
    rates = {}
    for n_thr in (1, 2, 4, 8, 16, 32, 64):
        counts = [0] * n_thr
        stop = time.perf_counter() + 2.0
        bad = []

        def worker(k):
            pk, props = streams[k % len(streams)]
            dst = np.zeros(65536, np.uint8)
            dl, sl = C.c_size_t(0), C.c_size_t(0)
            while time.perf_counter() < stop:
                dl.value, sl.value = 65536, len(pk)
                r = L.LzmaUncompress(dst.ctypes.data, C.byref(dl), pk.ctypes.data, C.byref(sl), props.ctypes.data, 5)
                if r != 0 or dl.value != 65536:
                    bad.append((k, r))
                    return
                counts[k] += 1
            if not np.array_equal(dst, corpus[(k % 64) * 65536:(k % 64 + 1) * 65536]):
                bad.append((k, "bytes"))

        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(k,)) for k in range(n_thr)]
        [t.start() for t in th]
        [t.join() for t in th]
        dt = time.perf_counter() - t0
        rates[str(n_thr)] = dict(calls_per_s=sum(counts) / dt, MBps=sum(counts) * 65536 / dt / 1e6, ok=not bad)
    out["LzmaUncompress_64KiB_concurrent_callers"] = rates
    out["speedup_64_callers_over_1"] = rates["64"]["calls_per_s"] / rates["1"]["calls_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
