"""GPU parity tests proper: everything goes through the C-ABI of lib/liblzb200.so on a real B200 and is
compared with (a) the committed golden vectors the reference produced and (b) the oracle on freshly
seeded inputs -- bit-exact bytes and identical (SRes, status, destLen, srcLen).  At larger sizes the
checks are size-independent properties (per-unit CRC fused on the GPU, round trips of reference-encoded
data, digest of digests)."""
import zlib

import numpy as np
import pytest

import cases
import lzb200
from lzb200 import records as R
from tools.corpus import add_x86_callsites, make_corpus

pytestmark = pytest.mark.gpu


class Gpu:
    """refbind-shaped adapter over the product's C-ABI (lzb200.sdk)."""
    kind = "gpu"

    def lzma_decode(self, packed, props, dest_cap, finish_mode=0):
        return lzb200.sdk.lzma_decode(packed, props, dest_cap, finish_mode)

    def lzma2_decode(self, packed, prop, dest_cap, finish_mode=1):
        return lzb200.sdk.lzma2_decode(packed, prop, dest_cap, finish_mode)

    def xz_decode(self, data, dest_cap, finish_mode=0):
        return lzb200.sdk.xz_decode(data, dest_cap)

    def x86_convert(self, data, ip=0, state=0, encoding=0):
        return lzb200.sdk.x86_convert(data, ip, state, encoding)

    def delta(self, data, dist, encoding=0):
        return (lzb200.sdk.delta_encode if encoding else lzb200.sdk.delta_decode)(data, dist)[0]

    def bra_convert(self, fid, data, ip=0, encoding=0):
        return lzb200.sdk.bra_convert(fid, data, ip, encoding)

    def crc32(self, data):
        return lzb200.sdk.crc32(data)

    def crc64(self, data):
        return lzb200.sdk.crc64(data)


# Lanes per unit (lzb_kernels.cu): "auto" is what a user gets (two units per warp, or one when the launch
# has few units); 16 and 8 force the several-units-per-warp loops on the small test launches too.
GROUPS = ["auto", "16", "8"]


def _set_group(g):
    import os
    if g == "auto":
        os.environ.pop("LZB_GROUP", None)
    else:
        os.environ["LZB_GROUP"] = g


@pytest.fixture(scope="module", params=GROUPS)
def gpu(request):
    _set_group(request.param)
    yield Gpu()
    _set_group("auto")


@pytest.fixture(scope="module")
def ctx(gpu):
    c = lzb200.lib.Context(0)
    yield c
    c.close()


def same(a, b):
    for f in ("res", "status", "dest_len", "src_len"):
        assert a[f] == b[f], (f, {k: a[k] for k in a if k != "out"}, {k: b[k] for k in b if k != "out"})
    assert np.array_equal(a["out"][: a["dest_len"]], b["out"][: b["dest_len"]])


# ------------------------------------------------------------------ golden vectors (reference-made)
@pytest.mark.parametrize("kind", ["lzma", "lzma2"])
def test_golden_decodes(gpu, kind):
    blobs, _ = cases.golden()
    for c in cases.cases_of(kind):
        cases.check_decode_case(gpu, c, blobs)


def test_golden_xz(gpu):
    blobs, _ = cases.golden()
    for c in cases.cases_of("xz"):
        cases.check_xz_case(gpu, c, blobs)


@pytest.mark.parametrize("kind", ["x86", "delta", "crc"])
def test_golden_filters_and_checks(gpu, kind):
    blobs, _ = cases.golden()
    for c in cases.cases_of(kind):
        cases.check_filter_case(gpu, c, blobs)


def test_golden_branch_converters_and_delta_encode(gpu):
    cs, blobs = cases.bra_golden()
    for c in cs:
        cases.check_bra_case(gpu, c, blobs)


# ------------------------------------------------------------------ fresh inputs vs the oracle
def test_lzma_fuzz_against_oracle(gpu, ora):
    """bit flips / truncations / odd capacities on golden streams, compared with the oracle."""
    blobs, _ = cases.golden()
    rng = np.random.default_rng(77)
    seen = set()
    for c in [x for x in cases.cases_of("lzma") if x["mut"] is None and x["finish"] == 0 and x["cap"] > 0][:9]:
        packed, props, n = blobs[c["blob"]], bytes.fromhex(c["props"]), c["cap"]
        for _ in range(12):
            bad = packed.copy()
            if len(bad) > 8:
                bad[int(rng.integers(0, len(bad)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            cut = int(rng.integers(0, len(bad) + 1)) if rng.integers(0, 3) == 0 else len(bad)
            cap = int(rng.integers(0, n + 50)) if rng.integers(0, 3) == 0 else n
            fm = int(rng.integers(0, 2))
            a, b = ora.lzma_decode(bad[:cut], props, cap, fm), gpu.lzma_decode(bad[:cut], props, cap, fm)
            same(a, b)
            seen.add(a["res"])
    assert {0, 1, 6} <= seen


def test_batch_matches_per_stream_calls(gpu, ora, ctx):
    """LzmaUncompressBatch == n LzmaUncompress calls, including bad members that must not poison neighbours."""
    blobs, _ = cases.golden()
    good = [x for x in cases.cases_of("lzma") if x["mut"] is None and x["finish"] == 0]
    packed_list, caps, props = [], [], []
    rng = np.random.default_rng(5)
    for i in range(300):
        c = good[i % len(good)]
        p = blobs[c["blob"]]
        if i % 17 == 3 and len(p) > 30:
            p = p.copy()
            p[int(rng.integers(5, len(p)))] ^= 0x40
        if i % 29 == 7:
            p = p[: len(p) // 2]
        packed_list.append(p)
        caps.append(c["cap"])
        props.append(bytes.fromhex(c["props"]))
    out = lzb200.sdk.lzma_uncompress_batch(ctx, packed_list, b"".join(props), caps)
    for i, d in enumerate(out):
        want = ora.lzma_decode(packed_list[i], props[i], caps[i], 0)
        assert (d["res"], d["dest_len"], d["src_len"]) == (want["res"], want["dest_len"], want["src_len"]), i
        assert np.array_equal(d["out"][: d["dest_len"]], want["out"][: want["dest_len"]]), i


def test_many_units_digest_of_digests(gpu, ctx, ora):
    """4096 raw LZMA units through Lzb_DecodeUnitsHost with the CRC-32 fused on the GPU: every unit must verify
    against the digest of its plain text, and the XOR of all GPU digests equals the host's."""
    blobs, _ = cases.golden()
    src = [x for x in cases.cases_of("lzma") if x["mut"] is None and x["finish"] == 0 and x["cap"] >= 5000
           and x["props"] == "5d00000100" and x["res"] == 0 and x["out_crc"] == x["plain_crc"]]
    assert src
    n = 4096
    lens = np.array([len(blobs[src[i % len(src)]["blob"]]) for i in range(n)], dtype=np.uint32)
    caps = np.array([src[i % len(src)]["cap"] for i in range(n)], dtype=np.uint32)
    soff = np.concatenate(([0], np.cumsum((lens.astype(np.uint64) + 3) & ~np.uint64(3))[:-1])).astype(np.uint64)
    doff = np.concatenate(([0], np.cumsum(caps.astype(np.uint64))[:-1])).astype(np.uint64)
    packed = lzb200.sdk.pinned_empty(int(soff[-1] + lens[-1] + 16))
    for i in range(n):
        packed[int(soff[i]):int(soff[i]) + int(lens[i])] = blobs[src[i % len(src)]["blob"]]
    out = lzb200.sdk.pinned_empty(int(doff[-1] + caps[-1]))
    units = lzb200.sdk.raw_lzma_units(soff, lens, doff, caps, bytes.fromhex("5d00000100"))
    units["check"], units["flags"] = R.CHECK_CRC32, R.UNIT_VERIFY
    units["expect"] = [src[i % len(src)]["plain_crc"] for i in range(n)]
    want_src = np.array([src[i % len(src)]["src_len"] for i in range(n)], dtype=np.uint32)   # excludes an end mark
    want_dst = np.array([src[i % len(src)]["dest_len"] for i in range(n)], dtype=np.uint32)  # < cap when a mark ends it
    import os
    # both host paths: sliced launches, and ONE launch fed and drained while it runs (LzbFlow)
    for flow in ("0", "1"):
        os.environ["LZB_FLOW"] = flow
        try:
            out[:] = 0
            res = lzb200.sdk.decode_units_host(ctx, packed, out, units)
        finally:
            os.environ.pop("LZB_FLOW", None)
        assert (res["res"] == 0).all(), (flow, np.unique(res["res"], return_counts=True))
        assert (res["dst_done"] == want_dst).all() and (res["src_used"] == want_src).all(), flow
        assert np.bitwise_xor.reduce(res["check"]) == np.bitwise_xor.reduce(units["expect"]), flow
        for k in (0, 1234, n - 1):
            assert zlib.crc32(out[int(doff[k]):int(doff[k]) + int(caps[k])].tobytes()) == src[k % len(src)]["plain_crc"], (flow, k)
    lzb200.sdk.pinned_free_all()


def test_lzma2_parallel_units_and_limits(gpu, ora):
    blobs, _ = cases.golden()
    blob = blobs["lzma2_multi"]
    prop = [c["prop"] for c in cases.cases_of("lzma2") if c["blob"] == "lzma2_multi"][0]
    total = 360000
    rng = np.random.default_rng(9)
    for cap in (total, total + 1, total - 1, 90000, 89999, 90001, 180000, 1, 0):
        for fm in (0, 1):
            same(ora.lzma2_decode(blob, prop, cap, fm), gpu.lzma2_decode(blob, prop, cap, fm))
    for _ in range(25):
        bad = blob.copy()
        bad[int(rng.integers(0, len(bad)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        cut = int(rng.integers(0, len(bad) + 1)) if rng.integers(0, 4) == 0 else len(bad)
        for fm in (0, 1):
            same(ora.lzma2_decode(bad[:cut], prop, total, fm), gpu.lzma2_decode(bad[:cut], prop, total, fm))


@pytest.mark.needref
def test_fresh_reference_encoded_round_trips(gpu, ref, ora):
    """new data, encoded by the reference encoder here, decoded on the GPU: LZMA, LZMA2 and xz."""
    corpus = add_x86_callsites(make_corpus(3 << 20, seed=4242), seed=4)
    packed, props = ref.lzma_encode(corpus[:200000], level=5, dict_size=1 << 16)
    d = gpu.lzma_decode(packed, props, 200000, 0)
    assert d["res"] == 0 and np.array_equal(d["out"], corpus[:200000])
    parts, prop = [], 0
    for k in range(3):
        pk, prop = ref.lzma2_encode(corpus[k << 20:(k + 1) << 20], level=1, dict_size=1 << 20)
        parts.append(pk if k == 2 else pk[:-1])
    d = gpu.lzma2_decode(np.concatenate(parts), prop, 3 << 20, 1)
    assert (d["res"], d["status"], d["dest_len"]) == (0, 1, 3 << 20) and np.array_equal(d["out"], corpus)
    # xz: 4 MiB-class blocks would take long on the slow reference encoder; 512 KiB blocks, BCJ + CRC64
    pk_list, lens, ulen, checks = [], [], [], []
    for b0 in range(0, 2 << 20, 1 << 19):
        blk = corpus[b0:b0 + (1 << 19)]
        pk, prop = ref.lzma2_encode(ref.x86_filter_whole(blk, 0, 1), level=1, dict_size=1 << 20)
        pk_list.append(pk); lens.append(len(pk)); ulen.append(len(blk)); checks.append(ref.crc64(blk))
    offs = np.concatenate(([0], np.cumsum(lens)[:-1]))
    xz = ref.xz_assemble(np.concatenate(pk_list), offs, lens, ulen, checks, 4, 4, 0, prop)
    a, b = ref.xz_decode(xz, (2 << 20) + 8), gpu.xz_decode(xz, (2 << 20) + 8)
    assert (b["res"], b["status"], b["dest_len"], b["src_len"], b["num_blocks"], b["finished"]) == \
           (a["res"], a["status"], a["dest_len"], a["src_len"], a["num_blocks"], a["finished"])
    assert np.array_equal(b["out"][: 2 << 20], corpus[: 2 << 20])


def test_filters_and_checks_large(gpu, ora):
    data = add_x86_callsites(make_corpus(3_000_001, seed=11), seed=12, every=31)
    assert gpu.crc32(data) == ora.crc32(data) == zlib.crc32(data.tobytes())
    assert gpu.crc64(data) == ora.crc64(data)
    assert lzb200.sdk.crc32_update(0x1234, data[:70001]) == (zlib.crc32(data[:70001].tobytes(), 0x1234 ^ 0xFFFFFFFF) ^ 0xFFFFFFFF)
    for enc in (0, 1):
        a, b = ora.x86_convert(data, 0, 0, enc), gpu.x86_convert(data, 0, 0, enc)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
    # large buffers are cut into segments converted by different warps (lzb_host.cu): carried state, odd sizes, a
    # start offset, and data that is nothing but opcode candidates (no place to cut)
    dense = np.full(600_000, 0xE8, np.uint8)
    for buf, ip, st, enc in ((data[:1_000_003], 0x40001000, 5, 0), (data[7:2_500_000], 12345, 3, 1), (dense, 0, 1, 0),
                             (np.concatenate((dense[:300_000], data[:400_000])), 0x7FFFFFF0, 0, 0)):
        a, b = ora.x86_convert(buf, ip, st, enc), gpu.x86_convert(buf, ip, st, enc)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:], (len(buf), ip, st, enc, a[1:], b[1:])
    for dist in (1, 2, 3, 4, 33, 256):
        assert np.array_equal(ora.delta(data, dist, 0), gpu.delta(data, dist))
    # carried history across two Delta_Decode calls == one call
    one = ora.delta(data[:100000], 5, 0)
    first, st = lzb200.sdk.delta_decode(data[:40001], 5)
    second, _ = lzb200.sdk.delta_decode(data[40001:100000], 5, st)
    assert np.array_equal(np.concatenate((first, second)), one)


def test_x86_convert_small_buffers_with_carried_state(gpu, ora):
    """Chunked x86_Convert use: a non-zero incoming state still reaches candidates at positions 0..2 (Bra86.c:18-19),
    which matters when the buffer is so short that the lanes' start positions fall there (5..200 bytes, all 8 states,
    both directions)."""
    rng = np.random.default_rng(5)
    bad = 0
    for n in list(range(5, 150)) + [199, 200]:
        for rep in range(6):
            data = rng.integers(0, 256, n, dtype=np.uint8)
            hot = rng.random(n) < 0.35
            data[hot] = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF], np.uint8), int(hot.sum()))
            st = (rep + n) & 7
            enc = rep & 1
            a, b = ora.x86_convert(data, 0x1000 * rep, st, enc), gpu.x86_convert(data, 0x1000 * rep, st, enc)
            bad += not (np.array_equal(a[0], b[0]) and a[1:] == b[1:])
    assert bad == 0


def test_concurrent_one_call_decodes_are_combined_and_exact(gpu, ora):
    """LzmaUncompress / LzmaDecode from many host threads at once: the calls share launches (lzb_host.cu, "callers that
    do not batch") and every caller still gets exactly its own (SRes, destLen, srcLen) and bytes -- good streams,
    truncated ones and short outputs mixed."""
    import ctypes as C
    import threading
    L = lzb200.lib.load()
    blobs, cs = cases.golden()
    good = [c for c in cs if c["kind"] == "lzma" and c["mut"] is None and c["finish"] == 0 and c["cap"] > 0][:24]
    jobs = []
    for i, c in enumerate(good):
        pk, props = blobs[c["blob"]], np.frombuffer(bytes.fromhex(c["props"]), np.uint8).copy()
        variant = i % 3
        src = pk if variant != 1 else pk[: max(6, len(pk) // 2)]
        cap = c["cap"] if variant != 2 else max(1, c["cap"] // 3)
        jobs.append((np.ascontiguousarray(src), props, cap, ora.lzma_decode(src, props, cap, 0)))
    errors = []

    def worker(k):
        src, props, cap, want = jobs[k % len(jobs)]
        for _ in range(6):
            dst = np.zeros(cap + 8, np.uint8)
            dl, sl = C.c_size_t(cap), C.c_size_t(len(src))
            r = L.LzmaUncompress(dst.ctypes.data, C.byref(dl), src.ctypes.data, C.byref(sl), props.ctypes.data, 5)
            if (r, dl.value, sl.value) != (want["res"], want["dest_len"], want["src_len"]) or \
                    not np.array_equal(dst[: dl.value], want["out"][: dl.value]):
                errors.append((k, r, dl.value, sl.value))
    th = [threading.Thread(target=worker, args=(k,)) for k in range(48)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors[:5]


def test_one_table_over_several_contexts(gpu, ora):
    """Lzb_DecodeUnitsHostMulti: one unit table cut over several contexts (here: several contexts of GPU 0 -- the
    partition, the per-range arena windows and the threading are what is under test) == the single-context call."""
    import ctypes as C
    L = lzb200.lib.load()
    L.Lzb_DecodeUnitsHostMulti.restype = C.c_int
    L.Lzb_DecodeUnitsHostMulti.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
    blobs, cs = cases.golden()
    good = [c for c in cs if c["kind"] == "lzma" and c["mut"] is None and c["finish"] == 0 and c["cap"] > 0]
    props0 = good[0]["props"]
    good = [c for c in good if c["props"] == props0] * 9
    n = len(good)
    plen = np.array([len(blobs[c["blob"]]) for c in good], dtype=np.uint64)
    ulen = np.array([c["cap"] for c in good], dtype=np.uint64)
    soff = np.concatenate(([0], np.cumsum((plen + 15) & ~np.uint64(15))[:-1])).astype(np.uint64)
    doff = np.concatenate(([0], np.cumsum(ulen)[:-1])).astype(np.uint64)
    packed = np.zeros(int(soff[-1] + plen[-1]) + 64, np.uint8)
    for i, c in enumerate(good):
        packed[int(soff[i]):int(soff[i]) + int(plen[i])] = blobs[c["blob"]]
    units = lzb200.sdk.raw_lzma_units(soff, plen.astype(np.uint32), doff, ulen.astype(np.uint32), bytes.fromhex(props0))
    ctxs = [lzb200.lib.Context(0) for _ in range(3)]
    one = np.zeros(int(doff[-1] + ulen[-1]) + 64, np.uint8)
    r1 = lzb200.sdk.decode_units_host(ctxs[0], packed, one, units)
    for k in (2, 3):
        out = np.zeros_like(one)
        res = np.zeros(n, dtype=R.RESULT_DTYPE)
        arr = (C.c_void_p * k)(*[c.h for c in ctxs[:k]])
        rc = L.Lzb_DecodeUnitsHostMulti(arr, k, packed.ctypes.data, packed.size, out.ctypes.data, out.size, units.ctypes.data, res.ctypes.data, n)
        assert rc == 0 and np.array_equal(res, r1)
        for j in range(n):                      # (bytes behind a unit's dst_done are not part of the contract)
            o, m = int(doff[j]), int(res["dst_done"][j])
            assert np.array_equal(out[o:o + m], one[o:o + m]), (k, j)
    for c in ctxs:
        c.close()


def test_golden_lzma86_one_call(gpu):
    """Lzma86_Decode / Lzma86_GetUnpackSize (Lzma86Dec.c) vs the reference's answers: tests/golden/lzma86.json."""
    import json
    import os
    from tools.make_golden_lzma86 import bind, decode, mutate
    here = os.path.dirname(os.path.abspath(__file__))
    cs = json.load(open(os.path.join(here, "golden", "lzma86.json")))["cases"]
    z = np.load(os.path.join(here, "golden", "lzma86.npz"))
    L = bind(lzb200.lib.load())
    assert len(cs) > 60
    for c in cs:
        m = tuple(c["mut"]) if c["mut"] else None
        got = decode(L, mutate(z[c["blob"]].tobytes(), m), c["cap"])
        want = {k: c[k] for k in got}
        assert got == want, (c["blob"], c["mut"], c["cap"], got, want)
        if c["mut"] is None and c["res"] == 0 and c["dest_len"] == c["size"]:
            assert c["out_crc"] == c["plain_crc"]


@pytest.mark.needref
def test_bcj2_decode_stand_alone(gpu, ref):
    """Bcj2_Decode (Bcj2.c:28-127) vs the reference on streams split by the fixture encoder: whole, short output,
    truncated side streams, corrupted range-coder stream."""
    import ctypes as C
    from oracle import refbind
    from tools import sz_writer as W
    R = C.CDLL(refbind.REF_SO)
    P = lzb200.lib.load()
    for L in (R, P):
        L.Bcj2_Decode.restype = C.c_int
        L.Bcj2_Decode.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                  C.c_void_p, C.c_size_t]
    data = add_x86_callsites(make_corpus(150000, seed=77), seed=3).tobytes()
    m, c, j, r = W.bcj2_encode(data)
    variants = [(m, c, j, r, len(data)), (m, c, j, r, len(data) // 2), (m, c, j, r, 0), (m, c[:-4], j, r, len(data)), (m, c, j[:8], r, len(data)),
                (m, c, j, r[: len(r) // 2], len(data)), (m, c, j, r[:3], len(data)), (m[:-10], c, j, r, len(data)), (m, c, j, bytes([r[0] ^ 0x40]) + r[1:], len(data)),
                (m, c, j, r[:20] + bytes([r[20] ^ 0x08]) + r[21:], len(data))]
    for k, (a0, a1, a2, a3, n) in enumerate(variants):
        o1, o2 = np.zeros(n + 8, np.uint8), np.zeros(n + 8, np.uint8)
        r1 = R.Bcj2_Decode(a0, len(a0), a1, len(a1), a2, len(a2), a3, len(a3), o1.ctypes.data, n)
        r2 = P.Bcj2_Decode(a0, len(a0), a1, len(a1), a2, len(a2), a3, len(a3), o2.ctypes.data, n)
        assert r1 == r2, (k, r1, r2)
        if r1 == 0:
            assert np.array_equal(o1[:n], o2[:n]), k
    assert R.Bcj2_Decode(m, len(m), c, len(c), j, len(j), r, len(r), o1.ctypes.data, len(data)) == 0 and o1[: len(data)].tobytes() == data
