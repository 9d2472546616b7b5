#!/usr/bin/env python
"""Generate tests/golden/vectors.npz + vectors.json from the UNMODIFIED reference (oracle/_ref/libref.so).

The reference ships no golden vectors, and /root/reference does not exist on the GPU box,
so the vectors are produced HERE by running the reference itself (encoder for the inputs, decoder for
the expected (SRes, status, destLen, srcLen) and a CRC-32 of the expected bytes) and committed.
Re-run:  python tools/make_golden.py      (deterministic: seeded corpus, fixed mutation recipes)

Case kinds
  lzma   : LzmaDecode(packed[mutated], props, cap, finish)
  lzma2  : Lzma2Dec_AllocateProbs/Init/DecodeToDic(packed[mutated], prop, cap, finish)
  xz     : XzUnpacker_Code loop over a whole buffer
  x86    : x86_Convert(data, ip, state, encoding)          delta : Delta_Decode   crc : CrcCalc / Crc64Calc
Mutations: ("cut", n) keeps the first n bytes; ("xor", pos, mask) flips bits; None leaves the blob alone.
"""
from __future__ import annotations

import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refbind  # noqa: E402
from tools.corpus import add_x86_callsites, entropy_sweep_mix, make_corpus  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")


def mutate(blob: np.ndarray, m):
    if m is None:
        return blob
    if m[0] == "cut":
        return blob[: m[1]]
    out = blob.copy()
    out[m[1]] ^= np.uint8(m[2])
    return out


def main():
    refbind.build(ref=True)
    ref = refbind.ref()
    rng = np.random.default_rng(20261018)
    corpus = make_corpus(2 << 20, seed=1234)
    blobs, cases = {}, []

    def add_blob(name, arr):
        blobs[name] = np.ascontiguousarray(arr, dtype=np.uint8)

    def expect(d, n_out=None):
        n = d["dest_len"] if n_out is None else n_out
        return dict(res=int(d["res"]), status=int(d["status"]), dest_len=int(d["dest_len"]), src_len=int(d["src_len"]),
                    out_crc=zlib.crc32(d["out"][:n].tobytes()))

    # ---- raw LZMA streams
    specs = [(65536, 3, 0, 2, 0, 1 << 16, 5), (65536, 3, 0, 2, 1, 1 << 16, 5), (1, 3, 0, 2, 0, 1 << 16, 5),
             (0, 3, 0, 2, 1, 1 << 16, 5), (5000, 4, 0, 0, 1, 4096, 1), (20000, 0, 4, 4, 0, 8192, 3),
             (30000, 1, 2, 1, 1, 4096, 0), (40000, 3, 1, 0, 1, 8192, 6), (50000, 8, 0, 2, 0, 1 << 15, 2)]
    for i, (n, lc, lp, pb, em, ds, lvl) in enumerate(specs):
        off = int(rng.integers(0, len(corpus) - max(n, 1)))
        plain = corpus[off:off + n]
        packed, props = ref.lzma_encode(plain, level=lvl, dict_size=ds, lc=lc, lp=lp, pb=pb, end_mark=em)
        name = f"lzma{i}"
        add_blob(name, packed)
        muts = [None, ("cut", 0), ("cut", 4), ("cut", 5), ("cut", max(len(packed) - 1, 0)), ("cut", len(packed) // 2),
                ("cut", max(len(packed) - 20, 0)), ("cut", max(len(packed) - 21, 0))]
        for _ in range(6):
            if len(packed) > 6:
                muts.append(("xor", int(rng.integers(0, len(packed))), int(1 << rng.integers(0, 8))))
        muts.append(("xor", 0, 1))
        for m in muts:
            if m and m[0] == "xor" and m[1] >= len(packed):
                continue
            blob = mutate(packed, m)
            caps = sorted({n, n + 7, n // 2, 0}) if m is None else [n]
            for cap in caps:
                for fm in (0, 1):
                    d = ref.lzma_decode(blob, props, cap, fm)
                    cases.append(dict(kind="lzma", blob=name, mut=m, props=bytes(props).hex(), cap=cap, finish=fm,
                                      plain_crc=zlib.crc32(plain.tobytes()), **expect(d)))
    # entropy sweep
    for t in (0.0, 0.5, 1.0):
        plain = make_corpus(60000, seed=int(t * 10) + 40, mix=entropy_sweep_mix(t))
        packed, props = ref.lzma_encode(plain, level=5, dict_size=1 << 16)
        name = f"sweep{int(t * 10)}"
        add_blob(name, packed)
        d = ref.lzma_decode(packed, props, len(plain), 0)
        cases.append(dict(kind="lzma", blob=name, mut=None, props=bytes(props).hex(), cap=len(plain), finish=0,
                          plain_crc=zlib.crc32(plain.tobytes()), **expect(d)))

    # ---- LZMA2
    l2specs = [(300000, 1, 1 << 18, "corpus"), (70000, 3, 1 << 12, "corpus"), (1, 0, 4096, "corpus"),
               (120000, 1, 1 << 16, "random")]
    for i, (n, lvl, ds, src) in enumerate(l2specs):
        plain = corpus[1000:1000 + n] if src == "corpus" else rng.integers(0, 256, size=n, dtype=np.uint8)
        packed, prop = ref.lzma2_encode(plain, level=lvl, dict_size=ds)
        name = f"lzma2_{i}"
        add_blob(name, packed)
        muts = [None, ("cut", len(packed) - 1), ("cut", 0), ("cut", 2), ("cut", 6), ("cut", len(packed) // 3)]
        for _ in range(6):
            muts.append(("xor", int(rng.integers(0, len(packed))), int(1 << rng.integers(0, 8))))
        for m in muts:
            blob = mutate(packed, m)
            for cap in ([n, n + 5, n // 2] if m is None else [n]):
                for fm in (0, 1):
                    d = ref.lzma2_decode(blob, prop, cap, fm)
                    cases.append(dict(kind="lzma2", blob=name, mut=m, prop=int(prop), cap=cap, finish=fm,
                                      plain_crc=zlib.crc32(plain.tobytes()), **expect(d)))
    # multi-unit stream: independent units, EOS stripped from all but the last
    parts, plains = [], []
    for k in range(4):
        plain = corpus[200000 + k * 90000: 200000 + (k + 1) * 90000]
        pk, prop = ref.lzma2_encode(plain, level=1, dict_size=1 << 17)
        parts.append(pk if k == 3 else pk[:-1])
        plains.append(plain)
    multi = np.concatenate(parts)
    whole = np.concatenate(plains)
    add_blob("lzma2_multi", multi)
    for m in (None, ("cut", len(multi) - 1), ("cut", len(parts[0]) + 10), ("xor", len(parts[0]) + len(parts[1]) + 100, 4),
              ("xor", len(parts[0]), 0x20)):
        blob = mutate(multi, m)
        for cap in ([len(whole), len(whole) + 9, 200000, 180000] if m is None else [len(whole)]):
            for fm in (0, 1):
                d = ref.lzma2_decode(blob, prop, cap, fm)
                cases.append(dict(kind="lzma2", blob="lzma2_multi", mut=m, prop=int(prop), cap=cap, finish=fm,
                                  plain_crc=zlib.crc32(whole.tobytes()), **expect(d)))

    # ---- xz containers
    data = add_x86_callsites(corpus[:400000], seed=8)
    for j, (ct, fid, fp) in enumerate(((4, 4, 0), (1, 0, 0), (4, 3, 4), (0, 3, 256))):
        pk_list, lens, ulen, checks = [], [], [], []
        prop = 0
        for b0 in range(0, len(data), 131072):
            blk = data[b0:b0 + 131072]
            f = ref.x86_filter_whole(blk, 0, 1) if fid == 4 else (ref.delta(blk, fp, 1) if fid == 3 else blk)
            pk, prop = ref.lzma2_encode(f, level=1, dict_size=1 << 20)
            pk_list.append(pk); lens.append(len(pk)); ulen.append(len(blk))
            checks.append(ref.crc64(blk) if ct == 4 else (ref.crc32(blk) if ct == 1 else 0))
        offs = np.concatenate(([0], np.cumsum(lens)[:-1]))
        xz = ref.xz_assemble(np.concatenate(pk_list), offs, lens, ulen, checks, ct, fid, fp if fid == 3 else 0, prop)
        name = f"xz{j}"
        add_blob(name, xz)
        muts = [None, ("cut", len(xz) - 1), ("cut", len(xz) // 2), ("cut", 12), ("cut", 13), ("xor", 20, 0x10),
                ("xor", len(xz) // 2, 0x10), ("xor", len(xz) - 14, 0x10), ("xor", len(xz) - 3, 0x10), ("xor", 7, 0x10)]
        for m in muts:
            d = ref.xz_decode(mutate(xz, m), len(data) + 16)
            e = expect(d) if m is None else dict(res=int(d["res"]), status=int(d["status"]), dest_len=-1, src_len=-1, out_crc=0)
            cases.append(dict(kind="xz", blob=name, mut=m, cap=len(data) + 16, num_blocks=int(d["num_blocks"]),
                              finished=bool(d["finished"]), plain_crc=zlib.crc32(data.tobytes()), **e))

    # ---- filters and checks on raw data
    raw = add_x86_callsites(corpus[500000:540000], seed=5, every=23)
    add_blob("raw", raw)
    soup = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF, 0x12], dtype=np.uint8), size=3000)
    add_blob("soup", soup)
    for name, arr in (("raw", raw), ("soup", soup)):
        for ip, st, enc in ((0, 0, 0), (0, 0, 1), (0x12345678, 0, 0), (4096, 5, 0), (0, 3, 1)):
            out, n, st2 = ref.x86_convert(arr, ip, st, enc)
            cases.append(dict(kind="x86", blob=name, ip=ip, state=st, enc=enc, processed=n, state_out=st2,
                              out_crc=zlib.crc32(out.tobytes())))
        for dist in (1, 2, 3, 4, 17, 255, 256):
            cases.append(dict(kind="delta", blob=name, dist=dist, out_crc=zlib.crc32(ref.delta(arr, dist, 0).tobytes())))
        for cut in (0, 1, 3, 4, 5, 63, 64, 65, 1000, len(arr)):
            seg = arr[:cut]
            cases.append(dict(kind="crc", blob=name, n=cut, crc32=ref.crc32(seg), crc64=ref.crc64(seg)))

    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, "vectors.npz"), **blobs)
    with open(os.path.join(OUT_DIR, "vectors.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden.py", reference="LZMA SDK 9.20 (yurket/lzma-sdk-zlibLike), gcc -O2 -D_7ZIP_ST",
                       cases=cases), f, separators=(",", ":"))
    total = sum(b.nbytes for b in blobs.values())
    print(f"{len(cases)} cases, {len(blobs)} blobs, {total} blob bytes")


if __name__ == "__main__":
    main()
