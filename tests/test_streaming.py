"""The zlib-like resumable contract: LzmaDec_DecodeToBuf / Lzma2Dec_DecodeToBuf fed and
drained in arbitrary chunk sizes must give what the reference gives -- same bytes, same final
(SRes, status, destLen, srcLen) -- for valid, truncated and corrupted streams.

CPU part : oracle vs reference (pins the oracle's streaming), and the device code's resume logic under
           emulation (tests/streamdrive.EmuStream) vs the oracle.
GPU part : the real C-ABI objects on the B200 vs the oracle.
"""
import numpy as np
import pytest

import cases
from streamdrive import EmuStream, GpuStream

CHUNKINGS = ((1, 1 << 20), (7, 13), (1000, 777), (1 << 20, 1), (19, 4096), (64, 64), (1 << 20, 1 << 20), (5, 5), (20, 21))


def same(a, b):
    for f in ("res", "status", "dest_len", "src_len"):
        assert a[f] == b[f], (f, {k: a[k] for k in a if k != "out"}, {k: b[k] for k in b if k != "out"})
    assert np.array_equal(a["out"][: a["dest_len"]], b["out"][: b["dest_len"]])


def lzma_streams(limit=None):
    blobs, _ = cases.golden()
    out = []
    for c in cases.cases_of("lzma"):
        if c["mut"] is None and c["finish"] == 0 and c["res"] == 0 and c["out_crc"] == c["plain_crc"] and c["dest_len"] == c["cap"]:
            out.append((blobs[c["blob"]], bytes.fromhex(c["props"]), c["cap"]))
    return out[:limit] if limit else out


def lzma2_streams():
    blobs, _ = cases.golden()
    seen, out = set(), []
    for c in cases.cases_of("lzma2"):
        if c["mut"] is None and c["blob"] not in seen and c["res"] == 0 and c["finish"] == 1 and c["status"] == 1:
            seen.add(c["blob"])
            out.append((blobs[c["blob"]], c["prop"], c["dest_len"]))
    return out


def run_matrix(impl, want, streams2=True, heavy=False):
    rng = np.random.default_rng(31)
    for packed, props, n in lzma_streams():
        for ic, oc in CHUNKINGS:
            if not heavy and n > 6000 and min(ic, oc) < 19:
                continue
            for fm in (0, 1):
                same(want.lzma_decode_tobuf(packed, props, n, ic, oc, fm), impl.lzma_decode_tobuf(packed, props, n, ic, oc, fm))
        if len(packed) > 40:
            for _ in range(3):
                bad = packed.copy()
                bad[int(rng.integers(5, len(bad)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
                cut = int(rng.integers(0, len(bad)))
                for blob in (bad, packed[:cut]):
                    # exact also on failure: the commit points inside a failing call are part of the contract
                    same(want.lzma_decode_tobuf(blob, props, n, 997, 4096, 1), impl.lzma_decode_tobuf(blob, props, n, 997, 4096, 1))
                    same(want.lzma_decode_tobuf(blob, props, n, 64, 300, 0), impl.lzma_decode_tobuf(blob, props, n, 64, 300, 0))
    if not streams2:
        return
    for packed, prop, n in lzma2_streams():
        for ic, oc in ((1 << 20, 1 << 20), (1000, 777), (4096, 1 << 20), (1 << 20, 5000), (3, 100000) if n < 100000 else (333, 65536)):
            for fm in (0, 1):
                same(want.lzma2_decode_tobuf(packed, prop, n, ic, oc, fm), impl.lzma2_decode_tobuf(packed, prop, n, ic, oc, fm))
        same(want.lzma2_decode_tobuf(packed[:-1], prop, n, 5000, 5000, 1), impl.lzma2_decode_tobuf(packed[:-1], prop, n, 5000, 5000, 1))
        same(want.lzma2_decode_tobuf(packed, prop, n // 2, 5000, 3000, 0), impl.lzma2_decode_tobuf(packed, prop, n // 2, 5000, 3000, 0))
        for _ in range(3):
            bad = packed.copy()
            bad[int(rng.integers(0, len(bad)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            for ic, oc, fm in ((997, 4096, 1), (4096, 50000, 0)):
                same(want.lzma2_decode_tobuf(bad, prop, n, ic, oc, fm), impl.lzma2_decode_tobuf(bad, prop, n, ic, oc, fm))


@pytest.mark.needref
def test_oracle_streaming_pinned_to_reference(ref, ora):
    run_matrix(ora, ref)


@pytest.mark.parametrize("form", ["lean", "ahead"])
def test_kernel_resume_logic_under_emulation(ora, form):
    import emubind
    emubind.AHEAD = 1 if form == "ahead" else 0
    try:
        run_matrix(EmuStream(), ora)
    finally:
        emubind.AHEAD = 0


@pytest.mark.gpu
def test_gpu_stateful_objects(ora):
    run_matrix(GpuStream(), ora)


@pytest.mark.gpu
def test_gpu_lzma2_whole_buffer_goes_parallel_and_matches(ora):
    """Lzma2Dec_DecodeToDic on a fresh object with a multi-unit buffer takes the parallel path; result and
    object state must be those of the sequential decoder."""
    import ctypes as C
    from streamdrive import CLzma2Dec, _alloc
    import lzb200
    g = GpuStream()
    L = g.L
    L.Lzma2Dec_AllocateProbs.restype = C.c_int
    L.Lzma2Dec_AllocateProbs.argtypes = [C.c_void_p, C.c_uint8, C.c_void_p]
    L.Lzma2Dec_DecodeToDic.restype = C.c_int
    L.Lzma2Dec_DecodeToDic.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_int, C.POINTER(C.c_int)]
    L.LzmaDec_FreeProbs.argtypes = [C.c_void_p, C.c_void_p]
    blobs, _ = cases.golden()
    prop = [c["prop"] for c in cases.cases_of("lzma2") if c["blob"] == "lzma2_multi"][0]
    one = blobs["lzma2_multi"][:-1]
    blob = np.concatenate([one] * 4 + [np.zeros(1, np.uint8)])        # 16 units, > 1 MiB packed? no: force the path
    total = 360000 * 4
    want = ora.lzma2_decode(blob, prop, total, 1)
    d = CLzma2Dec()
    assert L.Lzma2Dec_AllocateProbs(C.byref(d), prop, C.byref(_alloc)) == 0
    out = np.zeros(total, np.uint8)
    d.decoder.dic, d.decoder.dicBufSize = out.ctypes.data, total        # 7zDec.c:185-186
    L.Lzma2Dec_Init(C.byref(d))
    sl, st = C.c_size_t(len(blob)), C.c_int(0)
    r = L.Lzma2Dec_DecodeToDic(C.byref(d), total, blob.ctypes.data, C.byref(sl), 1, C.byref(st))
    assert (r, st.value, d.decoder.dicPos, sl.value) == (want["res"], want["status"], want["dest_len"], want["src_len"])
    assert np.array_equal(out, want["out"])
    L.LzmaDec_FreeProbs(C.byref(d.decoder), C.byref(_alloc))


@pytest.mark.gpu
@pytest.mark.needref
def test_gpu_config1_shape_single_stream_with_ring_dictionary(ref):
    """BASELINE config 1 in small: one long stream through LzmaDec_DecodeToBuf (the LzmaUtil loop), dictionary
    much smaller than the stream so the ring wraps many times; fed 64 KiB in / 64 KiB out like LzmaUtil.c:55-106."""
    from tools.corpus import make_corpus
    plain = make_corpus(3 << 20, seed=41)
    packed, props = ref.lzma_encode(plain, level=1, dict_size=1 << 18, end_mark=1)
    g = GpuStream()
    for ic, oc in ((1 << 16, 1 << 16), (100000, 1 << 20)):
        a = ref.lzma_decode_tobuf(packed, props, len(plain), ic, oc, 0)
        b = g.lzma_decode_tobuf(packed, props, len(plain), ic, oc, 0)
        same(a, b)
        assert np.array_equal(b["out"][: len(plain)], plain)


def ring_wrap_streams(ref, count=40):
    """A match whose length equals its distance and that starts exactly where the ring dictionary wraps
    (dicPos == 0, source = the last L bytes of the ring): the byte the next matched literal is compared with is the
    FIRST byte this copy writes (LzmaDec.c:178: dic[dicPos - rep0] after the copy)."""
    rng = np.random.default_rng(77)
    out = []
    for k in range(count):
        L = int(rng.integers(2, 273))
        head = rng.integers(0, 256, 4096, dtype=np.uint8)
        tail = rng.integers(0, 256, int(rng.integers(8, 600)), dtype=np.uint8)
        plain = np.concatenate((head, head[4096 - L:], tail))
        packed, props = ref.lzma_encode(plain, level=int(rng.integers(1, 10)), dict_size=4096, end_mark=int(k & 1))
        out.append((plain, packed, props))
    return out


@pytest.mark.needref
@pytest.mark.parametrize("form", ["lean", "ahead"])
def test_match_of_len_eq_distance_at_the_ring_wrap_under_emulation(ref, form):
    import emubind
    emubind.AHEAD = 1 if form == "ahead" else 0
    try:
        for plain, packed, props in ring_wrap_streams(ref):
            for ic, oc in ((1 << 20, 1 << 20), (4096, 4096), (997, 333)):
                a = ref.lzma_decode_tobuf(packed, props, len(plain), ic, oc, 0)
                b = EmuStream().lzma_decode_tobuf(packed, props, len(plain), ic, oc, 0)
                same(a, b)
                assert np.array_equal(b["out"][: len(plain)], plain)
    finally:
        emubind.AHEAD = 0


@pytest.mark.gpu
@pytest.mark.needref
def test_gpu_match_of_len_eq_distance_at_the_ring_wrap(ref):
    g = GpuStream()
    for plain, packed, props in ring_wrap_streams(ref, 24):
        for ic, oc in ((1 << 20, 1 << 20), (997, 333)):
            a = ref.lzma_decode_tobuf(packed, props, len(plain), ic, oc, 0)
            b = g.lzma_decode_tobuf(packed, props, len(plain), ic, oc, 0)
            same(a, b)
            assert np.array_equal(b["out"][: len(plain)], plain)


@pytest.mark.gpu
@pytest.mark.needref
def test_gpu_config1_at_baseline_size(ref):
    """BASELINE configs[0] at size: ONE 64 MiB stream, lc3 lp0 pb2, 16 MiB dictionary (the ring wraps three times),
    through LzmaDec_Allocate + LzmaDec_DecodeToBuf fed 64 KiB in / 64 KiB out like LzmaUtil.c:55-106 -- same
    (SRes, status, destLen, srcLen) and bytes as the reference's own loop.  One stream = one warp: ~15 s on a B200."""
    import zlib
    from tools.corpus import make_corpus
    plain = make_corpus(64 << 20, seed=4321)
    packed, props = ref.lzma_encode(plain, level=1, dict_size=16 << 20, end_mark=1)
    a = ref.lzma_decode_tobuf(packed, props, len(plain), 1 << 16, 1 << 16, 0)
    b = GpuStream().lzma_decode_tobuf(packed, props, len(plain), 1 << 16, 1 << 16, 0)
    for f in ("res", "status", "dest_len", "src_len"):
        assert a[f] == b[f], f
    assert zlib.crc32(b["out"][: len(plain)].tobytes()) == zlib.crc32(plain.tobytes())
