"""Differential fuzz of the device decode logic on the CPU: the same generator as tools/fuzz_gpu.py (random data shapes x
random LZMA properties, encoded by the reference encoder; pristine, bit-flipped, truncated, short output), decoded by
the host build of csrc/lzb_core.cuh (tests/emu, both forms of the symbol decoder) and by the unmodified reference.
Meant to run under AddressSanitizer + UBSan with arenas that end exactly where the contract says they may:

    LZB_EMU_SANITIZE=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \\
        ASAN_OPTIONS=detect_leaks=0 python tools/fuzz_emu.py [--rounds 20] [--seed 1]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools.fuzz_gpu import shape  # noqa: E402


def fuzz_x86(a, ref, emu, emubind):
    """x86_Convert: both converters of lzb_post.cuh (per-lane segments = lean form, tile scan = look-ahead form) on buffers
    full of opcode candidates and fill bytes, every carried state, both directions, unaligned starts"""
    rng = np.random.default_rng(a.seed)
    total = bad = 0
    for it in range(a.rounds * 500):
        n = int(rng.choice([5, 6, 7, 9, 15, 16, 17, 20, 31, 32, 33, 47, 48, 49, 64, 100, 257, 1000, 5000]))
        kind = int(rng.integers(0, 4))
        if kind == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            d = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF, 0x12, 0x80], dtype=np.uint8), n, p=[.25, .15, .25, .2, .1, .05])
        elif kind == 2:
            d = rng.choice(np.array([0xE8, 0x00, 0xFF], dtype=np.uint8), n)
        else:
            d = rng.integers(0, 256, n, dtype=np.uint8)
            for k in range(0, n - 5, int(rng.integers(5, 40))):
                d[k] = 0xE8
                d[k + 4] = 0 if rng.integers(0, 2) else 0xFF
        ip, st, enc = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 8)), int(rng.integers(0, 2))
        w = ref.x86_convert(d, ip, st, enc)
        for form in (0, 1):
            emubind.AHEAD = form
            g = emu.x86_convert(d, ip, st, enc)
            total += 1
            if not (np.array_equal(w[0], g[0]) and w[1] == g[1] and w[2] == g[2]):
                bad += 1
                print(f"MISMATCH x86 n {n} kind {kind} form {form} ip {ip} state {st} enc {enc}: ref {w[1:]} emu {g[1:]}", flush=True)
    emubind.AHEAD = 0
    print("FUZZ emu x86", "FAILED" if bad else "OK", total, "conversions", bad, "mismatches", "(sanitizers on)" if emubind.SANITIZE else "")
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--x86", action="store_true", help="fuzz the x86 BCJ converters instead of the decoder")
    a = ap.parse_args()
    import emubind
    from oracle import refbind
    ref = refbind.ref()
    emu = emubind.Emu()
    emubind.lib()
    if a.x86:
        return fuzz_x86(a, ref, emu, emubind)
    rng = np.random.default_rng(a.seed)
    total = bad = 0
    for rd in range(a.rounds):
        lc, lp = int(rng.integers(0, 9)), int(rng.integers(0, 5))
        if lc + lp > 8:
            lp = 0
        pb = int(rng.integers(0, 5))
        dict_size = int(rng.choice([1 << 12, 3 << 11, 1 << 14, 1 << 16]))
        level = int(rng.choice([0, 1, 5]))
        end_mark = int(rng.integers(0, 2))
        lzma2 = bool(rng.integers(0, 3) == 0)
        for k in range(16):
            n = int(rng.choice([0, 1, 2, 7, 100, 1000, 5000, 20000, 70000]))
            data = shape(rng, n) if n else np.zeros(0, np.uint8)
            if lzma2:
                pk, props = ref.lzma2_encode(data, level=level, dict_size=max(dict_size, 1 << 12))
            else:
                pk, props = ref.lzma_encode(data, level=level, dict_size=dict_size, lc=lc, lp=lp, pb=pb, fb=int(rng.choice([5, 32, 273])), end_mark=end_mark)
            for variant in range(4):
                blob, cap, fm = pk.copy(), n, int(rng.integers(0, 2))
                if variant == 1 and len(blob) > 6:
                    blob[int(rng.integers(1 if lzma2 else 5, len(blob)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
                elif variant == 2:
                    blob = blob[: int(rng.integers(0, len(blob) + 1))]
                elif variant == 3:
                    cap = int(rng.integers(0, n + 20))
                w = ref.lzma2_decode(blob, props, cap, fm) if lzma2 else ref.lzma_decode(blob, props, cap, fm)
                for form in (0, 1):
                    emubind.AHEAD = form
                    g = emu.lzma2_decode(blob, props, cap, fm) if lzma2 else emu.lzma_decode(blob, props, cap, fm)
                    total += 1
                    ok = (g["res"], g["status"], g["dest_len"], g["src_len"]) == (w["res"], w["status"], w["dest_len"], w["src_len"]) and \
                        np.array_equal(g["out"][: g["dest_len"]], w["out"][: w["dest_len"]])
                    if not ok:
                        bad += 1
                        print(f"MISMATCH round {rd} stream {k} variant {variant} form {form}: lzma2 {lzma2} lc{lc} lp{lp} pb{pb} dict {dict_size} "
                              f"emu {(g['res'], g['status'], g['dest_len'], g['src_len'])} ref {(w['res'], w['status'], w['dest_len'], w['src_len'])}", flush=True)
        print(f"round {rd}: {'lzma2' if lzma2 else f'lc{lc} lp{lp} pb{pb}'} dict {dict_size} level {level}: ok so far {total - bad}/{total}", flush=True)
    emubind.AHEAD = 0
    print("FUZZ emu", "FAILED" if bad else "OK", total, "decodes", bad, "mismatches", "(sanitizers on)" if emubind.SANITIZE else "")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
