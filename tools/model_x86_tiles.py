"""Python model of the G-lane tile walk of x86_convert_tiles (csrc/lzb_post.cuh) with explicit lanes, checked against the
unmodified reference's x86_Convert on random buffers (G = 1, 16, 32; every misalignment of the buffer start; all carried
states; both directions).  Validates the cross-lane logic that the single-lane host build of the device code cannot.

    python tools/model_x86_tiles.py
"""
import sys
import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refbind
ref = refbind.ref()

def is_fill(b): return b == 0 or b == 0xFF
def cand(data, pos, ip5, enc, hist, last):
    allow, topbit = 0x17, 0x6DB488
    gap = (pos - last) & 0xFFFFFFFF
    last = pos
    if gap > 3: hist = 0
    else:
        hist = (hist << (gap - 1)) & 7
        if hist:
            b = data[pos + 4 - ((topbit >> (3 * hist)) & 7)]
            if not ((allow >> hist) & 1) or is_fill(b):
                return pos + 1, ((hist << 1) & 7) | 1, last, None
    if not is_fill(data[pos + 4]):
        return pos + 1, ((hist << 1) & 7) | 1, last, None
    v = (int(data[pos+4]) << 24) | (int(data[pos+3]) << 16) | (int(data[pos+2]) << 8) | int(data[pos+1])
    here = (ip5 + pos) & 0xFFFFFFFF
    while True:
        out = (here + v) & 0xFFFFFFFF if enc else (v - here) & 0xFFFFFFFF
        if hist == 0: break
        sh = ((topbit >> (3 * hist)) & 7) * 8
        if not is_fill((out >> (24 - sh)) & 0xFF): break
        v = out ^ ((1 << (32 - sh)) - 1)
    return pos + 5, hist, last, out

def model(data, ip, state, enc, G, mis):
    data = data.copy(); orig = data.copy(); size = len(data)
    if size < 5: return data, 0, state
    stop = size - 4
    hist, last, pos = state & 7, 0xFFFFFFFF, 0
    ip5 = (ip + 5) & 0xFFFFFFFF
    tile = G * 16
    tb = -mis
    def load(at):
        return [int(orig_now[at + j]) if 0 <= at + j < size else 0 for j in range(16)]
    while tb < stop:
        orig_now = data          # loads see current memory (like the GPU); marks below pos are ignored
        masks = []
        for lane in range(G):
            at = tb + lane * 16
            b = load(at)
            m = 0
            for j in range(16):
                if (b[j] & 0xFE) == 0xE8: m |= 1 << j
            if at + 16 > stop: m = 0 if at >= stop else m & ((1 << (stop - at)) - 1)
            masks.append(m)
        while True:
            best = None
            for lane in range(G):
                at = tb + lane * 16
                rel = pos - at
                mm = masks[lane] if rel <= 0 else (0 if rel >= 16 else masks[lane] & (0xFFFF << rel) & 0xFFFF)
                if mm:
                    best = at + (mm & -mm).bit_length() - 1
                    break
            if best is None: break
            pos, hist, last, out = cand(data, best, ip5, enc, hist, last)
            if out is not None:
                c = best
                data[c+4] = (0 - ((out >> 24) & 1)) & 0xFF; data[c+3] = (out >> 16) & 0xFF; data[c+2] = (out >> 8) & 0xFF; data[c+1] = out & 0xFF
        tb += tile
    if pos < stop: pos = stop
    gap = (pos - last) & 0xFFFFFFFF
    state = 0 if gap > 3 else (hist << (gap - 1)) & 7
    return data, pos, state

rng = np.random.default_rng(5)
bad = 0
for it in range(4000):
    n = int(rng.choice([5, 17, 100, 300, 511, 512, 513, 700, 1500, 3000]))
    kind = int(rng.integers(0, 3))
    if kind == 0: d = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == 1: d = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF, 0x12, 0x80], dtype=np.uint8), n, p=[.25, .15, .25, .2, .1, .05])
    else: d = rng.choice(np.array([0xE8, 0x00, 0xFF], dtype=np.uint8), n)
    ip = int(rng.integers(0, 1 << 32)); st = int(rng.integers(0, 8)); enc = int(rng.integers(0, 2))
    G = int(rng.choice([1, 16, 32])); mis = int(rng.integers(0, 16))
    w = ref.x86_convert(d, ip, st, enc)
    g = model(d, ip, st, enc, G, mis)
    if not (np.array_equal(w[0], g[0]) and w[1] == g[1] and w[2] == g[2]):
        bad += 1
        if bad < 5: print("MISMATCH", n, kind, G, mis, w[1:], g[1:])
print("model", bad, "mismatches")
