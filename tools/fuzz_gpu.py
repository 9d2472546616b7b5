"""Differential fuzz on a GPU: random data shapes x random LZMA properties, encoded by the reference encoder,
decoded by liblzb200.so (one batched pass per property set) and by the unmodified reference; pristine and
damaged.  Every (SRes, destLen, srcLen) and every output byte must agree.  Not part of pytest (minutes of GPU
time); run it when the decoder changes:

    python tools/fuzz_gpu.py [--rounds 40] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.corpus import add_x86_callsites, make_corpus  # noqa: E402


def shape(rng, n):
    kind = int(rng.integers(0, 6))
    if kind == 0:
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == 1:
        return np.full(n, int(rng.integers(0, 256)), dtype=np.uint8)
    if kind == 2:
        return make_corpus(n, seed=int(rng.integers(0, 1 << 30)))
    if kind == 3:
        period = int(rng.integers(1, 300))
        return np.tile(rng.integers(0, 256, period, dtype=np.uint8), n // period + 1)[:n].copy()
    if kind == 4:
        return add_x86_callsites(make_corpus(max(n, 64), seed=int(rng.integers(0, 1 << 30))), seed=int(rng.integers(0, 99)))[:n].copy()
    return (np.arange(n) * int(rng.integers(1, 7)) // int(rng.integers(1, 9))).astype(np.uint8)


def fuzz_lzma2(a, ref, lzb200, ctx):
    """multi-unit LZMA2 streams (independently encoded units, EOS stripped from all but the last) through
    Lzma2Decode, which cuts them at the dictionary resets and decodes the units in parallel"""
    rng = np.random.default_rng(a.seed)
    total = bad = 0
    for rd in range(a.rounds):
        parts, plains = [], []
        prop = 0
        nunits = int(rng.integers(1, 7))
        ds = int(rng.choice([1 << 12, 1 << 16, 1 << 20]))
        for k in range(nunits):
            n = int(rng.choice([1, 50, 3000, 70000, 200000, 2200000]))
            data = shape(rng, n)
            pk, prop = ref.lzma2_encode(data, level=int(rng.choice([0, 1, 5])), dict_size=ds)
            parts.append(pk if k == nunits - 1 else pk[:-1])
            plains.append(data)
        blob = np.concatenate(parts)
        total_out = sum(len(p) for p in plains)
        for variant in range(6):
            b, cap, fm = blob.copy(), total_out, int(rng.integers(0, 2))
            if variant == 1:
                b[int(rng.integers(0, len(b)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            elif variant == 2:
                b = b[: int(rng.integers(0, len(b) + 1))]
            elif variant == 3:
                cap = int(rng.integers(0, total_out + 10))
            elif variant == 4:
                b = b[:-1]
            g = lzb200.sdk.lzma2_decode(b, prop, cap, fm, ctx)
            w = ref.lzma2_decode(b, prop, cap, fm)
            total += 1
            ok = (g["res"], g["status"], g["dest_len"], g["src_len"]) == (w["res"], w["status"], w["dest_len"], w["src_len"]) and \
                np.array_equal(g["out"][: g["dest_len"]], w["out"][: w["dest_len"]])
            if not ok:
                bad += 1
                print(f"MISMATCH lzma2 round {rd} variant {variant}: units {nunits} dict {ds} gpu {(g['res'], g['status'], g['dest_len'], g['src_len'])} "
                      f"ref {(w['res'], w['status'], w['dest_len'], w['src_len'])}", flush=True)
    print("FUZZ lzma2", "FAILED" if bad else "OK", total, "decodes", bad, "mismatches")
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--lzma2", action="store_true")
    a = ap.parse_args()
    import lzb200
    from oracle import refbind
    ref = refbind.ref()
    ctx = lzb200.lib.Context(0)
    if a.lzma2:
        return fuzz_lzma2(a, ref, lzb200, ctx)
    rng = np.random.default_rng(a.seed)
    total = bad = 0
    for rd in range(a.rounds):
        lc, lp = int(rng.integers(0, 9)), int(rng.integers(0, 5))
        if lc + lp > 8:
            lp = 0
        pb = int(rng.integers(0, 5))
        dict_size = int(rng.choice([1 << 12, 3 << 11, 1 << 14, 1 << 16, 1 << 18]))
        level = int(rng.choice([0, 1, 5]))
        end_mark = int(rng.integers(0, 2))
        packed_list, caps, fms, plains = [], [], [], []
        props = None
        for k in range(48):
            n = int(rng.choice([0, 1, 2, 7, 100, 1000, 5000, 20000, 70000]))
            data = shape(rng, n) if n else np.zeros(0, np.uint8)
            pk, props = ref.lzma_encode(data, level=level, dict_size=dict_size, lc=lc, lp=lp, pb=pb, fb=int(rng.choice([5, 32, 273])), end_mark=end_mark)
            for variant in range(3):
                blob = pk.copy()
                cap = n
                if variant == 1 and len(blob) > 6:
                    blob[int(rng.integers(5, len(blob)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
                if variant == 2:
                    if rng.integers(0, 2):
                        blob = blob[: int(rng.integers(0, len(blob) + 1))]
                    else:
                        cap = int(rng.integers(0, n + 20))
                packed_list.append(blob); caps.append(cap); plains.append(data)
        got = lzb200.sdk.lzma_uncompress_batch(ctx, packed_list, props.tobytes(), caps)
        for i, g in enumerate(got):
            w = ref.lzma_decode(packed_list[i], props, caps[i], 0)
            total += 1
            # LzmaUncompress == LzmaDecode with FINISH_ANY, status dropped
            ok = (g["res"], g["dest_len"], g["src_len"]) == (w["res"], w["dest_len"], w["src_len"]) and \
                np.array_equal(g["out"][: g["dest_len"]], w["out"][: w["dest_len"]])
            if not ok:
                bad += 1
                print(f"MISMATCH round {rd} stream {i}: lc{lc} lp{lp} pb{pb} dict {dict_size} gpu {(g['res'], g['dest_len'], g['src_len'])} "
                      f"ref {(w['res'], w['dest_len'], w['src_len'])}", flush=True)
        print(f"round {rd}: lc{lc} lp{lp} pb{pb} dict {dict_size} level {level} endmark {end_mark}: {len(got)} streams ok so far {total - bad}/{total}", flush=True)
    print("FUZZ", "FAILED" if bad else "OK", total, "decodes", bad, "mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
