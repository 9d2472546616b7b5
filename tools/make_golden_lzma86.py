"""Golden vectors for Lzma86_Decode (LZMA + x86 filter, one call): streams written by the reference's
Lzma86_Encode, decoded by the reference's Lzma86_Decode, pristine / truncated / corrupted / short output.

    python tools/make_golden_lzma86.py   ->  tests/golden/lzma86.json + lzma86.npz
"""
import ctypes as C
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from tools.corpus import add_x86_callsites, make_corpus  # noqa: E402


def bind(L):
    L.Lzma86_Decode.restype = C.c_int
    L.Lzma86_Decode.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t)]
    L.Lzma86_GetUnpackSize.restype = C.c_int
    L.Lzma86_GetUnpackSize.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
    return L


def decode(L, blob: bytes, cap: int):
    src = np.frombuffer(blob + b"\0" * 8, dtype=np.uint8).copy()
    dst = np.zeros(cap + 8, np.uint8)
    dl, sl, us = C.c_size_t(cap), C.c_size_t(len(blob)), C.c_uint64(0)
    g = L.Lzma86_GetUnpackSize(src.ctypes.data, len(blob), C.byref(us))
    r = L.Lzma86_Decode(dst.ctypes.data, C.byref(dl), src.ctypes.data, C.byref(sl))
    return dict(res=int(r), dest_len=int(dl.value), src_len=int(sl.value), size_res=int(g), size=int(us.value) if g == 0 else 0,
                out_crc=int(zlib.crc32(dst[:dl.value].tobytes())) if r == 0 else 0)


def mutate(b, m):
    if m is None:
        return b
    if m[0] == "cut":
        return b[:m[1]]
    a = bytearray(b)
    a[m[1]] ^= m[2]
    return bytes(a)


def main():
    from oracle import refbind
    L = bind(C.CDLL(refbind.REF_SO))
    L.Lzma86_Encode.restype = C.c_int
    L.Lzma86_Encode.argtypes = [C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_int]
    corpus = make_corpus(200000, seed=61)
    code = add_x86_callsites(corpus[:90000], seed=6)
    rng = np.random.default_rng(8)
    blobs, cases = {}, []
    for name, data, mode in (("text_nofilter", corpus[100000:160000], 0), ("code_filter", code, 1), ("code_auto", code, 2), ("tiny", corpus[:7], 1)):
        data = np.ascontiguousarray(data)
        dst = np.zeros(len(data) * 2 + 4096, np.uint8)
        dl = C.c_size_t(len(dst))
        r = L.Lzma86_Encode(dst.ctypes.data, C.byref(dl), data.ctypes.data, len(data), 1, 1 << 16, mode)
        assert r == 0, r
        blob = dst[:dl.value].tobytes()
        blobs[name] = np.frombuffer(blob, dtype=np.uint8)
        n = len(blob)
        muts = [None, ("cut", n - 1), ("cut", n // 2), ("cut", 13), ("cut", 14), ("cut", 5), ("xor", 0, 1), ("xor", 0, 2), ("xor", 1, 0x80), ("xor", 7, 1),
                ("xor", 14, 1), ("xor", n // 2, 4), ("xor", n - 2, 0x10)]
        muts += [("xor", int(rng.integers(14, n)), 1 << int(rng.integers(0, 8))) for _ in range(6)]
        for m in muts:
            for cap in ((len(data), len(data) + 9, len(data) // 2) if m is None else (len(data),)):
                cases.append(dict(blob=name, mut=list(m) if m else None, cap=cap, plain_crc=int(zlib.crc32(data.tobytes())),
                                  **decode(L, mutate(blob, m), cap)))
    np.savez_compressed(os.path.join(OUT, "lzma86.npz"), **blobs)
    json.dump(dict(generator="tools/make_golden_lzma86.py", cases=cases), open(os.path.join(OUT, "lzma86.json"), "w"), separators=(",", ":"))
    import collections
    print(len(cases), "cases; res:", dict(collections.Counter(c["res"] for c in cases)), "filter bytes:", {k: int(v[0]) for k, v in blobs.items()})


if __name__ == "__main__":
    main()
