"""The resumable xz contract (XzUnpacker_Code, MixCoder_Code, BraState_Code, XzCheck_*):
XzUnpacker_Code fed and drained in arbitrary chunk sizes must give, AT EVERY CALL, the reference's
(SRes, status, srcLen, destLen), and the same bytes, stream-finished flag and block count at the end.

Golden: tests/golden/xz_stream.json -- chunked drives of the unmodified reference (tools/make_golden_xz.py)
over reference-written and liblzma-written files (CRC-32 / CRC-64 / SHA-256 / no check; x86, Delta, none;
multi-block, multi-stream, empty), whole / exact / short output buffers, truncations and bit flips.

CPU part : the committed vectors still equal the live reference (needref); public struct sizes.
GPU part : liblzb200.so's XzUnpacker_Code through the same driver vs the golden traces (and vs the live
           reference on fresh chunkings when libref.so travelled with the snapshot).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import xzdrive
from tools.make_golden_xz import mutate

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("res", "status", "src_len", "dest_len", "calls", "trace_crc", "finished", "num_blocks", "out_crc")


def golden():
    j = json.load(open(os.path.join(HERE, "golden", "xz_stream.json")))
    a = np.load(os.path.join(HERE, "golden", "vectors.npz"))
    b = np.load(os.path.join(HERE, "golden", "xz_stream.npz"))
    blobs = {k: a[k].tobytes() for k in ("xz0", "xz1", "xz2", "xz3")}
    blobs.update({k: b[k].tobytes() for k in b.files})
    return j["cases"], blobs


def check_case(L, c, blobs):
    m = tuple(c["mut"]) if c["mut"] else None
    d = xzdrive.drive_xz(L, mutate(blobs[c["blob"]], m), c["cap"], c["in_chunk"], c["out_chunk"], c["finish"])
    got = xzdrive.summary(d)
    want = {k: c[k] for k in KEYS}
    assert got == want, (c["blob"], c["mut"], c["in_chunk"], c["out_chunk"], c["cap"], c["finish"], got, want, d["trace"][:12].tolist())


@pytest.mark.needref
def test_golden_equals_live_reference(ref):
    from oracle import refbind
    L = xzdrive.bind(C.CDLL(refbind.REF_SO))
    cases, blobs = golden()
    for c in cases[::3]:
        check_case(L, c, blobs)


def test_golden_is_meaningful():
    cases, blobs = golden()
    assert len(cases) > 200
    whole = [c for c in cases if c["mut"] is None and c["res"] == 0 and c["finished"]]
    assert whole and all(c["out_crc"] == c["plain_crc"] and c["dest_len"] == c["total"] for c in whole)
    assert {c["res"] for c in cases} >= {0, 1, 3, 16, 17}        # OK, DATA, CRC, ARCHIVE, NO_ARCHIVE
    assert any(c["calls"] > 1000 for c in cases)


# ------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def product():
    import lzb200
    return xzdrive.bind(lzb200.lib.load())


@pytest.mark.gpu
@pytest.mark.parametrize("blob", ["xz0", "xz1", "xz2", "xz3", "sha_x86", "sha_two_streams", "tiny_crc64", "empty",
                                  "small_x86_crc32", "f_arm", "f_armt", "f_ppc", "f_sparc", "f_ia64", "chain3"])
def test_product_matches_golden_traces(product, blob):
    cases, blobs = golden()
    mine = [c for c in cases if c["blob"] == blob]
    assert mine
    for c in mine:
        check_case(product, c, blobs)


@pytest.mark.gpu
@pytest.mark.needref
def test_product_matches_live_reference_on_fresh_chunkings(product, ref):
    from oracle import refbind
    R = xzdrive.bind(C.CDLL(refbind.REF_SO))
    _, blobs = golden()
    rng = np.random.default_rng(5)
    for name in ("xz0", "xz2", "sha_x86", "sha_two_streams", "small_x86_crc32"):
        xz = blobs[name]
        for _ in range(4):
            ic = int(rng.integers(500, len(xz) + 10))
            oc = int(rng.integers(500, 200000))
            cap = int(rng.integers(1000, 420000))
            fm = int(rng.integers(0, 2))
            blob = xz
            if rng.integers(0, 3) == 0:
                blob = mutate(xz, ("xor", int(rng.integers(0, len(xz))), 1 << int(rng.integers(0, 8))))
            a = xzdrive.drive_xz(R, blob, cap, ic, oc, fm)
            b = xzdrive.drive_xz(product, blob, cap, ic, oc, fm)
            assert xzdrive.summary(a) == xzdrive.summary(b), (name, ic, oc, cap, fm, xzdrive.summary(a), xzdrive.summary(b))
            assert np.array_equal(a["out"], b["out"])


@pytest.mark.gpu
def test_check_objects_and_sha256(product):
    """XzCheck_Init/Update/Final and Sha256_* (GPU-hashed payload) against hashlib / zlib."""
    import hashlib
    import zlib
    L = product
    rng = np.random.default_rng(11)
    data = rng.integers(0, 256, 300001, dtype=np.uint8).tobytes()
    for mode, size, want in ((1, 4, zlib.crc32(data).to_bytes(4, "little")), (10, 32, hashlib.sha256(data).digest())):
        chk = C.create_string_buffer(120)
        L.XzCheck_Init(chk, mode)
        pos = 0
        for step in (1, 63, 64, 65, 4096, 100000, 300001):
            L.XzCheck_Update(chk, data[pos:pos + step], C.c_size_t(min(step, len(data) - pos)))
            pos = min(len(data), pos + step)
        out = C.create_string_buffer(64)
        assert L.XzCheck_Final(chk, out) == 1
        assert out.raw[:size] == want, mode
    sha = C.create_string_buffer(104)
    L.Sha256_Init(sha)
    L.Sha256_Update(sha, data, C.c_size_t(len(data)))
    out = C.create_string_buffer(32)
    L.Sha256_Final(sha, out)
    assert out.raw == hashlib.sha256(data).digest()
    chk = C.create_string_buffer(120)
    L.XzCheck_Init(chk, 7)                                     # an id nobody verifies (Xz.c:84-86)
    assert L.XzCheck_Final(chk, out) == 0
