#!/usr/bin/env python
"""Symbol statistics of the bench corpora: what the serial chain of one unit is made of.

Analysis tool, CPU only (not on any product or test path): a plain LZMA symbol parser written from the format
(LzmaDec.c:153-410 is the reference's version of the same loop), run over a few of the bench's uniquely encoded
units; its output is checked against the corpus, so the counts are those of a correct decode.

  python tools/symbol_stats.py [--workload lzma64k|lzma1m_l1] [--units N]

Prints per-output-byte and per-symbol figures used in profiles/r01_l_symbol_stats.md.
"""
from __future__ import annotations

import argparse
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tools.corpus import make_corpus  # noqa: E402


class Rc:
    def __init__(self, buf):
        self.b, self.p = buf, 5
        self.range, self.code = 0xFFFFFFFF, int.from_bytes(bytes(buf[1:5]), "big")
        self.model_bits = self.direct_bits = self.norms = 0

    def norm(self):
        if self.range < (1 << 24):
            self.range = (self.range << 8) & 0xFFFFFFFF
            self.code = ((self.code << 8) | self.b[self.p]) & 0xFFFFFFFF
            self.p += 1
            self.norms += 1

    def bit(self, probs, i):
        self.norm()
        self.model_bits += 1
        p = probs[i]
        bound = (self.range >> 11) * p
        if self.code < bound:
            self.range = bound
            probs[i] = p + ((2048 - p) >> 5)
            return 0
        self.range -= bound
        self.code -= bound
        probs[i] = p - (p >> 5)
        return 1

    def direct(self, n):
        v = 0
        for _ in range(n):
            self.norm()
            self.direct_bits += 1
            self.range >>= 1
            b = 1 if self.code >= self.range else 0
            if b:
                self.code -= self.range
            v = (v << 1) | b
        return v

    def tree(self, probs, base, bits):
        m = 1
        for _ in range(bits):
            m = (m << 1) | self.bit(probs, base + m)
        return m - (1 << bits)

    def rtree(self, probs, base, bits):
        m, v = 1, 0
        for i in range(bits):
            b = self.bit(probs, base + m)
            m = (m << 1) | b
            v |= b << i
        return v


def parse(packed, lc, lp, pb, out_len):
    rc = Rc(packed)
    P = lambda n: [1024] * n
    is_match, is_rep, g0, g1, g2, rep0long = P(12 << 4), P(12), P(12), P(12), P(12), P(12 << 4)
    slot, spec, align = P(4 << 6), P(128), P(16)
    lit = P(0x300 << (lc + lp))
    lens = [dict(choice=P(2), low=P(16 << 3), mid=P(16 << 3), high=P(256)) for _ in range(2)]
    out = bytearray(out_len)
    pos, state, reps = 0, 0, [0, 0, 0, 0]
    c = Counter()
    lenhist = Counter()

    def length(L, ps):
        if rc.bit(L["choice"], 0) == 0:
            return rc.tree(L["low"], ps << 3, 3)
        if rc.bit(L["choice"], 1) == 0:
            return 8 + rc.tree(L["mid"], ps << 3, 3)
        return 16 + rc.tree(L["high"], 0, 8)

    while pos < out_len:
        ps = pos & ((1 << pb) - 1)
        before = rc.model_bits + rc.direct_bits
        if rc.bit(is_match, (state << 4) + ps) == 0:
            prev = out[pos - 1] if pos else 0
            base = 0x300 * (((pos & ((1 << lp) - 1)) << lc) + (prev >> (8 - lc)))
            sym = 1
            if state >= 7:
                mb = out[pos - reps[0] - 1]
                while sym < 0x100:
                    mbit = (mb >> 7) & 1
                    mb = (mb << 1) & 0xFF
                    b = rc.bit(lit, base + ((1 + mbit) << 8) + sym)
                    sym = (sym << 1) | b
                    if mbit != b:
                        break
                c["matched_literal"] += 1
            else:
                c["plain_literal"] += 1
            while sym < 0x100:
                sym = (sym << 1) | rc.bit(lit, base + sym)
            out[pos] = sym & 0xFF
            pos += 1
            state = 0 if state < 4 else (state - 3 if state < 10 else state - 6)
            c["bits_literal"] += rc.model_bits + rc.direct_bits - before
            continue
        if rc.bit(is_rep, state):
            if rc.bit(g0, state) == 0:
                if rc.bit(rep0long, (state << 4) + ps) == 0:
                    state = 9 if state < 7 else 11
                    out[pos] = out[pos - reps[0] - 1]
                    pos += 1
                    c["short_rep"] += 1
                    c["bits_match"] += rc.model_bits + rc.direct_bits - before
                    continue
            else:
                if rc.bit(g1, state) == 0:
                    d = reps[1]
                else:
                    if rc.bit(g2, state) == 0:
                        d = reps[2]
                    else:
                        d = reps[3]
                        reps[3] = reps[2]
                    reps[2] = reps[1]
                reps[1] = reps[0]
                reps[0] = d
            n = length(lens[1], ps) + 2
            state = 8 if state < 7 else 11
            c["rep"] += 1
        else:
            reps[3], reps[2], reps[1] = reps[2], reps[1], reps[0]
            n = length(lens[0], ps) + 2
            state = 7 if state < 7 else 10
            ls = min(n - 2, 3)
            s = rc.tree(slot, ls << 6, 6)
            if s < 4:
                d = s
            else:
                nb = (s >> 1) - 1
                d = (2 | (s & 1)) << nb
                if s < 14:
                    d += rc.rtree(spec, d - s, nb)
                else:
                    d += rc.direct(nb - 4) << 4
                    d += rc.rtree(align, 0, 4)
            reps[0] = d
            c["match"] += 1
        n = min(n, out_len - pos)
        lenhist[min(n, 33)] += 1
        src = pos - reps[0] - 1
        for i in range(n):
            out[pos + i] = out[src + i]
        pos += n
        c["match_bytes"] += n
        c["bits_match"] += rc.model_bits + rc.direct_bits - before
    c["model_bits"], c["direct_bits"], c["norms"] = rc.model_bits, rc.direct_bits, rc.norms
    return bytes(out), c, lenhist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lzma64k", choices=["lzma64k", "lzma1m_l1"])
    ap.add_argument("--units", type=int, default=8)
    a = ap.parse_args()
    from oracle import refbind
    ref = refbind.ref()
    unit, level, ds = (65536, 5, 1 << 16) if a.workload == "lzma64k" else (1 << 20, 1, 1 << 20)
    uniq = 2048 if a.workload == "lzma64k" else 128
    corpus = make_corpus(uniq * unit, seed=1234)
    pick = np.linspace(0, uniq - 1, a.units).astype(int)
    offs = (pick * unit).astype(np.uint64)
    lens = np.full(len(pick), unit, dtype=np.uint32)
    enc = ref.encode_batch_mt(0, corpus, offs, lens, level=level, dict_size=ds, lc=3, lp=0, pb=2, fb=32 if level == 5 else 32,
                              end_mark=0, threads=4)
    tot, hist = Counter(), Counter()
    for k in range(len(pick)):
        packed = bytes(enc["buf"][int(enc["off"][k]):int(enc["off"][k]) + int(enc["len"][k])])
        out, c, lh = parse(packed, 3, 0, 2, unit)
        assert out == corpus[int(offs[k]):int(offs[k]) + unit].tobytes(), "parser output differs from the corpus"
        tot += c
        hist += lh
    nbytes = unit * len(pick)
    nsym = tot["plain_literal"] + tot["matched_literal"] + tot["short_rep"] + tot["rep"] + tot["match"]
    nlit = tot["plain_literal"] + tot["matched_literal"]
    ncopy = tot["rep"] + tot["match"]
    print(f"workload {a.workload}: {len(pick)} units of {unit} bytes, level {level}")
    print(f"  symbols per output byte          {nsym / nbytes:.3f}")
    print(f"  modelled bits per output byte    {tot['model_bits'] / nbytes:.2f}   direct bits {tot['direct_bits'] / nbytes:.2f}   "
          f"normalisations {tot['norms'] / nbytes:.2f}")
    print(f"  symbol mix: plain literal {tot['plain_literal'] / nsym:.3f}  matched literal {tot['matched_literal'] / nsym:.3f}  "
          f"match {tot['match'] / nsym:.3f}  rep {tot['rep'] / nsym:.3f}  short rep {tot['short_rep'] / nsym:.3f}")
    print(f"  P(literal) per symbol            {nlit / nsym:.3f}")
    print(f"  bits per literal symbol          {tot['bits_literal'] / max(nlit, 1):.2f}")
    print(f"  bits per match/rep symbol        {tot['bits_match'] / max(ncopy + tot['short_rep'], 1):.2f}")
    print(f"  bytes per match/rep              {tot['match_bytes'] / max(ncopy, 1):.2f}   share of output from copies "
          f"{tot['match_bytes'] / nbytes:.3f}")
    le = sum(v for k, v in hist.items() if k <= 32) / max(sum(hist.values()), 1)
    print(f"  copies of at most 32 bytes       {le:.3f}")


if __name__ == "__main__":
    main()
