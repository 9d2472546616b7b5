"""Shared case runner: committed golden vectors (tests/golden/, produced by the reference through
tools/make_golden.py) replayed against any implementation that offers the refbind method names
(lzma_decode, lzma2_decode, xz_decode, x86_convert, delta, crc32, crc64)."""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_G = None


def golden():
    global _G
    if _G is None:
        blobs = dict(np.load(os.path.join(HERE, "golden", "vectors.npz")))
        meta = json.load(open(os.path.join(HERE, "golden", "vectors.json")))
        _G = (blobs, meta["cases"])
    return _G


def mutate(blob, m):
    if m is None:
        return blob
    if m[0] == "cut":
        return blob[: m[1]]
    out = blob.copy()
    out[m[1]] ^= np.uint8(m[2])
    return out


def cases_of(kind):
    blobs, cases = golden()
    return [c for c in cases if c["kind"] == kind]


def check_decode_case(impl, c, blobs):
    """Replay one lzma / lzma2 golden case and assert the reference's observable results."""
    blob = mutate(blobs[c["blob"]], c["mut"])
    if c["kind"] == "lzma":
        d = impl.lzma_decode(blob, bytes.fromhex(c["props"]), c["cap"], c["finish"])
    else:
        d = impl.lzma2_decode(blob, c["prop"], c["cap"], c["finish"])
    got = (d["res"], d["status"], d["dest_len"], d["src_len"])
    want = (c["res"], c["status"], c["dest_len"], c["src_len"])
    assert got == want, (c["kind"], c["blob"], c["mut"], c["cap"], c["finish"], got, want)
    assert zlib.crc32(d["out"][: d["dest_len"]].tobytes()) == c["out_crc"], (c["blob"], c["mut"], "bytes differ")


def check_xz_case(impl, c, blobs):
    blob = mutate(blobs[c["blob"]], c["mut"])
    d = impl.xz_decode(blob, c["cap"])
    assert d["res"] == c["res"], (c["blob"], c["mut"], d["res"], c["res"])
    if c["res"] == 0:   # the flag is only meaningful on success (the reference sets it before it checks the footer)
        assert bool(d["finished"]) == c["finished"], (c["blob"], c["mut"], "finished flag")
    if c["mut"] is None:
        assert (d["status"], d["dest_len"], d["src_len"], d["num_blocks"]) == \
               (c["status"], c["dest_len"], c["src_len"], c["num_blocks"])
        assert zlib.crc32(d["out"][: d["dest_len"]].tobytes()) == c["out_crc"] == c["plain_crc"]


def check_filter_case(impl, c, blobs):
    data = blobs[c["blob"]]
    if c["kind"] == "x86":
        out, n, st = impl.x86_convert(data, c["ip"], c["state"], c["enc"])
        assert (n, st) == (c["processed"], c["state_out"]), (c, n, st)
        assert zlib.crc32(np.asarray(out).tobytes()) == c["out_crc"], c
    elif c["kind"] == "delta":
        out = impl.delta(data, c["dist"], 0)
        assert zlib.crc32(np.asarray(out).tobytes()) == c["out_crc"], c
    else:
        seg = data[: c["n"]]
        assert impl.crc32(seg) == c["crc32"], c
        assert impl.crc64(seg) == c["crc64"], c


# ---- stateless branch converters + Delta_Encode (tests/golden/bra.json, tools/make_golden_bra.py)
_bra = None


def bra_golden():
    global _bra
    if _bra is None:
        import json
        j = json.load(open(os.path.join(HERE, "golden", "bra.json")))
        z = np.load(os.path.join(HERE, "golden", "bra.npz"))
        _bra = (j["cases"], {k: z[k] for k in z.files})
    return _bra


def check_bra_case(impl, c, blobs):
    data = blobs[c["blob"]][: c["n"]]
    if c["kind"] == "bra":
        out, done = impl.bra_convert(c["id"], data, c["ip"], c["enc"])
        assert done == c["processed"], (c, done)
    else:
        out = impl.delta(data, c["dist"], 1)
    assert zlib.crc32(np.asarray(out).tobytes()) == c["out_crc"], c
