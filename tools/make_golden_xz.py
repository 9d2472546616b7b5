"""Golden vectors for the RESUMABLE xz path: chunked XzUnpacker_Code drives of the unmodified reference
(oracle/_ref/libref.so) over xz files, every call's (SRes, status, srcLen, destLen) digested into a trace CRC.

    python tools/make_golden_xz.py        ->  tests/golden/xz_stream.json (+ xz_stream.npz for the new blobs)

Blobs xz0..xz3 come from tests/golden/vectors.npz (made by the reference's own writers, tools/make_golden.py);
the SHA-256 / multi-stream / tiny files are written by liblzma (python `lzma`), an independent encoder.
"""
import ctypes as C
import json
import lzma
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools.corpus import add_x86_callsites, make_corpus  # noqa: E402
import xzdrive  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def first_payload_end(xz: bytes) -> int:
    """Offset just past the payload of the first block, from the stream's index (backward read)."""
    n = len(xz)
    while n and xz[n - 1] == 0 and xz[n - 2:n] != b"YZ":
        n -= 1
    foot = xz[n - 12:n]
    flags = (foot[8] << 8) | foot[9]
    t = flags & 0xF
    csz = 0 if t == 0 else 4 << ((t - 1) // 3)
    idx_size = (int.from_bytes(foot[4:8], "little") + 1) * 4
    idx = xz[n - 12 - idx_size:n - 12]
    pos = 1

    def varint():
        nonlocal pos
        v = s = 0
        while True:
            b = idx[pos]; pos += 1
            v |= (b & 0x7F) << s; s += 7
            if not b & 0x80:
                return v
    nb = varint()
    if nb == 0:
        return 12
    unpadded = varint()
    return 12 + unpadded - csz


def new_blobs():
    corpus = make_corpus(700000, seed=77)
    code = add_x86_callsites(corpus[:300000], seed=3)
    x86_l2 = [{"id": lzma.FILTER_X86}, {"id": lzma.FILTER_LZMA2, "preset": 1, "dict_size": 1 << 20}]
    l2 = [{"id": lzma.FILTER_LZMA2, "preset": 1, "dict_size": 1 << 18}]
    blobs, plains = {}, {}
    blobs["sha_x86"] = lzma.compress(code.tobytes(), format=lzma.FORMAT_XZ, check=lzma.CHECK_SHA256, filters=x86_l2)
    plains["sha_x86"] = code.tobytes()
    a, b = corpus[300000:370000].tobytes(), corpus[400000:470001].tobytes()
    blobs["sha_two_streams"] = (lzma.compress(a, format=lzma.FORMAT_XZ, check=lzma.CHECK_SHA256, filters=l2) + b"\0" * 8 +
                                lzma.compress(b, format=lzma.FORMAT_XZ, check=lzma.CHECK_SHA256, filters=l2))
    plains["sha_two_streams"] = a + b
    t = corpus[500000:505000].tobytes()
    blobs["tiny_crc64"] = lzma.compress(t, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, filters=l2)
    plains["tiny_crc64"] = t
    blobs["empty"] = lzma.compress(b"", format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC32)
    plains["empty"] = b""
    d = add_x86_callsites(corpus[510000:530000], seed=9).tobytes()
    blobs["small_x86_crc32"] = lzma.compress(d, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC32, filters=x86_l2)
    plains["small_x86_crc32"] = d
    # the other branch converters and a three-filter chain (liblzma writes them, the reference reads them)
    from tools.make_golden_bra import synth
    for name, fid, arch in (("f_arm", lzma.FILTER_ARM, "arm"), ("f_armt", lzma.FILTER_ARMTHUMB, "armt"), ("f_ppc", lzma.FILTER_POWERPC, "ppc"),
                            ("f_sparc", lzma.FILTER_SPARC, "sparc"), ("f_ia64", lzma.FILTER_IA64, "ia64")):
        data = np.concatenate((synth(arch, 150000, seed=7), corpus[540000:560000])).tobytes()
        blobs[name] = lzma.compress(data, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64,
                                    filters=[{"id": fid}, {"id": lzma.FILTER_LZMA2, "preset": 1, "dict_size": 1 << 18}])
        plains[name] = data
    data = add_x86_callsites(corpus[560000:700000], seed=12).tobytes()
    blobs["chain3"] = lzma.compress(data, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC32,
                                    filters=[{"id": lzma.FILTER_DELTA, "dist": 4}, {"id": lzma.FILTER_X86},
                                             {"id": lzma.FILTER_LZMA2, "preset": 1, "dict_size": 1 << 18}])
    plains["chain3"] = data
    return blobs, plains


def mutate(b: bytes, m):
    if m is None:
        return b
    if m[0] == "cut":
        return b[:m[1]]
    a = bytearray(b)
    a[m[1]] ^= m[2]
    return bytes(a)


def patterns(name, xz, total):
    n = len(xz)
    big = 1 << 30
    pe = first_payload_end(xz)
    ps = [(n, big, total + 16, 0), (n, big, total, 0), (n, big, total, 1), (4096, big, total + 16, 0),
          (n, 5000, total + 16, 0), (7777, 3333, total + 16, 1), (n, big, total // 2, 0), (n, big, max(total - 1, 0), 1),
          (pe, big, total + 16, 0), (pe + 1, big, total + 16, 0)]
    if total > 131072:
        ps += [(n, big, 131072, 0), (n, big, 131072, 1), (pe, big, 131072, 0), (pe, 131072, total + 16, 0),
               (n, 131072, total + 16, 0), (n, 65536, total + 16, 0)]
    if n < 8000:
        ps += [(1, big, total + 16, 0), (n, 1, total + 16, 0), (3, 7, total + 16, 1)]
    return ps


def main():
    from oracle import refbind
    L = xzdrive.bind(C.CDLL(refbind.REF_SO))
    old = np.load(os.path.join(OUT, "vectors.npz"))
    vj = json.load(open(os.path.join(OUT, "vectors.json")))
    plain_crc = {c["blob"]: c["plain_crc"] for c in vj["cases"] if c["kind"] == "xz"}
    total_of = {c["blob"]: c["dest_len"] for c in vj["cases"] if c["kind"] == "xz" and c["mut"] is None}
    blobs, plains = new_blobs()
    items = [(k, old[k].tobytes(), total_of[k], plain_crc[k], "vectors.npz") for k in ("xz0", "xz1", "xz2", "xz3")]
    items += [(k, v, len(plains[k]), zlib.crc32(plains[k]), "xz_stream.npz") for k, v in blobs.items()]
    cases = []
    for name, xz, total, pcrc, where in items:
        n = len(xz)
        for (ic, oc, cap, fm) in patterns(name, xz, total):
            d = xzdrive.drive_xz(L, xz, cap, max(ic, 1), oc, fm)
            cases.append(dict(blob=name, where=where, mut=None, in_chunk=max(ic, 1), out_chunk=oc, cap=cap, finish=fm,
                              plain_crc=pcrc, total=total, **xzdrive.summary(d)))
        if name in ("empty",):
            continue
        muts = [("cut", n - 1), ("cut", n // 2), ("cut", 13), ("xor", 20, 0x10), ("xor", n // 2, 0x10),
                ("xor", n - 14, 0x10), ("xor", n - 3, 0x10), ("xor", 7, 0x10), ("xor", first_payload_end(xz) + 1, 0x01)]
        for m in muts:
            for (ic, oc) in ((n, 1 << 30), (4096, 50000)):
                d = xzdrive.drive_xz(L, mutate(xz, m), total + 16, ic, oc, 0)
                cases.append(dict(blob=name, where=where, mut=list(m), in_chunk=ic, out_chunk=oc, cap=total + 16, finish=0,
                                  plain_crc=pcrc, total=total, **xzdrive.summary(d)))
    np.savez_compressed(os.path.join(OUT, "xz_stream.npz"), **{k: np.frombuffer(v, dtype=np.uint8) for k, v in blobs.items()})
    with open(os.path.join(OUT, "xz_stream.json"), "w") as f:
        json.dump(dict(generator="tools/make_golden_xz.py", reference="LZMA SDK 9.20 (yurket/lzma-sdk-zlibLike), gcc -O2 -D_7ZIP_ST",
                       cases=cases), f, separators=(",", ":"))
    ok = sum(1 for c in cases if c["res"] == 0)
    print(f"{len(cases)} cases ({ok} with res 0), {sum(len(v) for v in blobs.values())} new blob bytes")


if __name__ == "__main__":
    main()
