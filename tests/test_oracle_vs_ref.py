"""Pin the CPU restatement (oracle/*.c) against the UNMODIFIED reference (oracle/_ref/libref.so).

The reference ships no tests or golden vectors, so the oracle is pinned by
differential testing: same seeded inputs through both, compare (SRes, status, destLen, srcLen)
and the produced bytes -- over valid streams, every finish mode, short/long destinations,
truncated and bit-flipped inputs, chunked streaming, LZMA2 units, xz containers, filters, CRCs.
"""
import numpy as np
import pytest

from tools.corpus import add_x86_callsites, entropy_sweep_mix, make_corpus

pytestmark = pytest.mark.needref

FIELDS = ("res", "status", "dest_len", "src_len")


def same(a, b, bytes_too=True):
    for f in FIELDS:
        assert a[f] == b[f], (f, {k: a[k] for k in FIELDS}, {k: b[k] for k in FIELDS})
    if bytes_too:
        n = a["dest_len"]
        assert np.array_equal(a["out"][:n], b["out"][:n])


@pytest.fixture(scope="module")
def corpus():
    return make_corpus(3 << 20, seed=1234)


@pytest.fixture(scope="module")
def streams(ref, corpus):
    """(plain, packed, props) over a spread of sizes / parameters / end-mark settings."""
    out = []
    rng = np.random.default_rng(5)
    for i, (n, lc, lp, pb, em, ds) in enumerate([
            (65536, 3, 0, 2, 0, 1 << 16), (65536, 3, 0, 2, 1, 1 << 16), (1, 3, 0, 2, 0, 1 << 16),
            (0, 3, 0, 2, 1, 1 << 16), (7, 0, 0, 0, 0, 4096), (5000, 4, 0, 0, 1, 4096),
            (20000, 0, 4, 4, 0, 8192), (30000, 1, 2, 1, 1, 1 << 12), (200000, 8, 0, 2, 0, 1 << 15),
            (100000, 2, 2, 3, 0, 1 << 20), (40000, 3, 1, 0, 1, 1 << 13)]):
        off = int(rng.integers(0, len(corpus) - max(n, 1)))
        plain = corpus[off:off + n].copy()
        packed, props = ref.lzma_encode(plain, level=int(rng.integers(0, 7)), dict_size=ds, lc=lc, lp=lp, pb=pb,
                                        end_mark=em)
        out.append((plain, packed, props))
    return out


def test_struct_sizes(ref):
    # SURVEY.md §8b ABI table
    assert [ref.sizeof(i) for i in range(4)] == [136, 168, 1792, 16]


def test_lzma_valid_all_modes(ref, ora, streams):
    for plain, packed, props in streams:
        n = len(plain)
        for cap in sorted({n, n + 1, n + 100, n // 2, max(n - 1, 0), 0, 1}):
            for fm in (0, 1):
                a = ref.lzma_decode(packed, props, cap, fm)
                b = ora.lzma_decode(packed, props, cap, fm)
                same(a, b)
                if cap == n and a["res"] == 0:
                    assert np.array_equal(b["out"][:n], plain)


def test_lzma_truncated(ref, ora, streams):
    for plain, packed, props in streams:
        n = len(plain)
        cuts = sorted({0, 1, 4, 5, 6, 10, 19, 20, 21, len(packed) // 2, len(packed) - 21, len(packed) - 20,
                       len(packed) - 6, len(packed) - 5, len(packed) - 1})
        for cut in cuts:
            if cut < 0 or cut > len(packed):
                continue
            for fm in (0, 1):
                same(ref.lzma_decode(packed[:cut], props, n, fm), ora.lzma_decode(packed[:cut], props, n, fm))


def test_lzma_corrupted(ref, ora, streams):
    rng = np.random.default_rng(11)
    outcomes = set()
    for plain, packed, props in streams:
        if len(packed) < 8:
            continue
        n = len(plain)
        for _ in range(40):
            bad = packed.copy()
            k = int(rng.integers(0, len(bad)))
            bad[k] ^= np.uint8(1 << int(rng.integers(0, 8)))
            if rng.integers(0, 4) == 0:
                bad[int(rng.integers(0, len(bad)))] = rng.integers(0, 256)
            for fm in (0, 1):
                a = ref.lzma_decode(bad, props, n, fm)
                b = ora.lzma_decode(bad, props, n, fm)
                same(a, b)
                outcomes.add(a["res"])
    assert 1 in outcomes  # the fuzz really produced data errors


def test_lzma_bad_props(ref, ora, streams):
    plain, packed, props = streams[0]
    for d in (225, 255):
        p = props.copy()
        p[0] = d
        same(ref.lzma_decode(packed, p, len(plain), 0), ora.lzma_decode(packed, p, len(plain), 0))
    # first stream byte must be zero
    bad = packed.copy()
    bad[0] = 1
    same(ref.lzma_decode(bad, props, len(plain), 0), ora.lzma_decode(bad, props, len(plain), 0))


def test_lzma_random_garbage(ref, ora):
    rng = np.random.default_rng(3)
    props = np.array([0x5D, 0, 0, 1, 0], dtype=np.uint8)
    for _ in range(60):
        n = int(rng.integers(5, 400))
        junk = rng.integers(0, 256, size=n, dtype=np.uint8)
        junk[0] = 0
        for fm in (0, 1):
            same(ref.lzma_decode(junk, props, 5000, fm), ora.lzma_decode(junk, props, 5000, fm))


def test_lzma_streaming_chunked(ref, ora, streams):
    for plain, packed, props in streams[:9]:
        n = len(plain)
        for ic, oc in ((1, 1 << 20), (7, 13), (1000, 777), (1 << 20, 1), (19, 4096), (64, 64)):
            if n > 70000 and min(ic, oc) < 7:
                continue
            for fm in (0, 1):
                a = ref.lzma_decode_tobuf(packed, props, n, ic, oc, fm)
                b = ora.lzma_decode_tobuf(packed, props, n, ic, oc, fm)
                same(a, b)


def test_lzma2_units(ref, ora, corpus):
    rng = np.random.default_rng(17)
    for n, lvl, ds in ((1 << 20, 1, 1 << 20), (300000, 5, 1 << 16), (1, 0, 4096), (70000, 3, 1 << 12),
                       (2 << 20, 0, 1 << 20)):
        off = int(rng.integers(0, len(corpus) - n))
        plain = corpus[off:off + n]
        packed, prop = ref.lzma2_encode(plain, level=lvl, dict_size=ds)
        for cap in (n, n + 5, n // 2):
            for fm in (0, 1):
                a = ref.lzma2_decode(packed, prop, cap, fm)
                b = ora.lzma2_decode(packed, prop, cap, fm)
                same(a, b)
        assert np.array_equal(ora.lzma2_decode(packed, prop, n, 1)["out"], plain)
        # without the end marker
        same(ref.lzma2_decode(packed[:-1], prop, n, 1), ora.lzma2_decode(packed[:-1], prop, n, 1))
        same(ref.lzma2_decode(packed[:-1], prop, n, 0), ora.lzma2_decode(packed[:-1], prop, n, 0))
        for cut in (0, 1, 2, 3, 5, 6, 11, len(packed) // 3, len(packed) - 2):
            same(ref.lzma2_decode(packed[:cut], prop, n, 1), ora.lzma2_decode(packed[:cut], prop, n, 1))
        for _ in range(25):
            bad = packed.copy()
            bad[int(rng.integers(0, len(bad)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            same(ref.lzma2_decode(bad, prop, n, 1), ora.lzma2_decode(bad, prop, n, 1))


def test_lzma2_incompressible_uses_stored_chunks(ref, ora):
    rng = np.random.default_rng(23)
    plain = rng.integers(0, 256, size=200000, dtype=np.uint8)
    packed, prop = ref.lzma2_encode(plain, level=1, dict_size=1 << 16)
    assert packed[0] == 1  # stored chunk with dictionary reset
    same(ref.lzma2_decode(packed, prop, len(plain), 1), ora.lzma2_decode(packed, prop, len(plain), 1))
    idx = ora.lzma2_index(packed)
    assert idx["n"] == 1 and idx["has_eos"] and idx["units"][0]["unpack_len"] == len(plain)


def test_lzma2_multi_unit_index(ref, ora, corpus):
    # independent units = one Lzma2Enc_Encode per slice, EOS stripped from all but the last
    parts, plains = [], []
    for i in range(5):
        plain = corpus[i * 150000:(i + 1) * 150000]
        packed, prop = ref.lzma2_encode(plain, level=1, dict_size=1 << 18)
        parts.append(packed if i == 4 else packed[:-1])
        plains.append(plain)
    blob = np.concatenate(parts)
    idx = ora.lzma2_index(blob)
    assert idx["n"] == 5 and idx["has_eos"] and idx["end_off"] == len(blob)
    assert [u["unpack_len"] for u in idx["units"]] == [150000] * 5
    whole = np.concatenate(plains)
    a = ref.lzma2_decode(blob, prop, len(whole), 1)
    b = ora.lzma2_decode(blob, prop, len(whole), 1)
    same(a, b)
    assert np.array_equal(b["out"], whole)
    # each unit decodes standalone
    for u, plain in zip(idx["units"], plains):
        seg = blob[u["src_off"]:u["src_off"] + u["src_len"]]
        r = ora.lzma2_decode(seg, prop, len(plain), 0)
        assert r["res"] == 0 and r["dest_len"] == len(plain) and np.array_equal(r["out"], plain)


def test_crc(ref, ora, corpus):
    assert ora.crc32(b"123456789") == 0xCBF43926 and ora.crc64(b"123456789") == 0x995DC9BBDF1939FA
    rng = np.random.default_rng(1)
    for _ in range(30):
        a = int(rng.integers(0, 1000))
        n = int(rng.integers(0, 70000))
        seg = corpus[a:a + n]
        assert ref.crc32(seg) == ora.crc32(seg)
        assert ref.crc64(seg) == ora.crc64(seg)


def test_x86_filter(ref, ora, corpus):
    rng = np.random.default_rng(2)
    for trial in range(40):
        n = int(rng.integers(0, 30000))
        a = int(rng.integers(0, 100000))
        raw = add_x86_callsites(corpus[a:a + n], seed=trial, every=int(rng.integers(6, 200)))
        ip = int(rng.integers(0, 1 << 32)) if trial % 3 == 0 else 0
        st = int(rng.integers(0, 8)) if trial % 5 == 0 else 0
        for enc in (0, 1):
            ra = ref.x86_convert(raw, ip, st, enc)
            oa = ora.x86_convert(raw, ip, st, enc)
            assert np.array_equal(ra[0], oa[0]) and ra[1:] == oa[1:]
        enc = ref.x86_filter_whole(raw, 0, 1)
        assert np.array_equal(ora.x86_filter_whole(enc, 0, 0), raw)
    # dense opcode soup stresses the neighbour-history logic
    soup = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF, 0x12], dtype=np.uint8), size=5000)
    for enc in (0, 1):
        ra = ref.x86_convert(soup, 0, 0, enc)
        oa = ora.x86_convert(soup, 0, 0, enc)
        assert np.array_equal(ra[0], oa[0]) and ra[1:] == oa[1:]


def test_delta_filter(ref, ora, corpus):
    rng = np.random.default_rng(4)
    for dist in (1, 2, 3, 4, 7, 16, 255, 256):
        n = int(rng.integers(0, 5000))
        seg = corpus[1000:1000 + n]
        assert np.array_equal(ref.delta(seg, dist, 0), ora.delta(seg, dist, 0))
        assert np.array_equal(ref.delta(seg, dist, 1), ora.delta(seg, dist, 1))
        assert np.array_equal(ora.delta(ora.delta(seg, dist, 1), dist, 0), seg)


def _make_xz(ref, data, block, check_type, filter_id, fparam, level=1, dict_size=1 << 20):
    blocks = [data[i:i + block] for i in range(0, len(data), block)]
    packed, lens, checks, ulen = [], [], [], []
    prop = 0
    for b in blocks:
        if filter_id == 4:
            f = ref.x86_filter_whole(b, fparam, 1)
        elif filter_id == 3:
            f = ref.delta(b, fparam, 1)
        else:
            f = b
        pk, prop = ref.lzma2_encode(f, level=level, dict_size=dict_size)
        packed.append(pk)
        lens.append(len(pk))
        ulen.append(len(b))
        checks.append(ref.crc64(b) if check_type == 4 else (ref.crc32(b) if check_type == 1 else 0))
    offs = np.concatenate(([0], np.cumsum(lens)[:-1]))
    # the x86 start offset is not written as a filter property here (fparam only meaningful for delta)
    return ref.xz_assemble(np.concatenate(packed), offs, lens, ulen, checks, check_type, filter_id,
                           fparam if filter_id == 3 else 0, prop)


def test_xz_container(ref, ora, corpus):
    data = add_x86_callsites(corpus[:900000], seed=8)
    for check_type, filter_id, fparam in ((4, 4, 0), (1, 0, 0), (4, 3, 4), (0, 3, 256), (1, 4, 0)):
        xz = _make_xz(ref, data, 262144, check_type, filter_id, fparam)
        a = ref.xz_decode(xz, len(data) + 10)
        b = ora.xz_decode(xz, len(data) + 10)
        assert a["res"] == 0 and a["finished"] and a["num_blocks"] == 4
        same(a, b)
        assert b["finished"] and b["num_blocks"] == 4
        assert np.array_equal(b["out"][:len(data)], data)
        idx = ora.xz_index(xz)
        assert idx["res"] == 0 and idx["check_type"] == check_type and len(idx["blocks"]) == 4
        assert sum(x["unpack_size"] for x in idx["blocks"]) == len(data)
        assert all(x["filter_id"] == filter_id for x in idx["blocks"])
        # corrupt payload / digest / index / footer: same SRes class
        rng = np.random.default_rng(check_type * 10 + filter_id)
        for pos in (20, len(xz) // 2, len(xz) - 40, len(xz) - 14, len(xz) - 3, 7, 9):
            bad = xz.copy()
            bad[pos] ^= 0x10
            ra = ref.xz_decode(bad, len(data) + 10)
            rb = ora.xz_decode(bad, len(data) + 10)
            assert ra["res"] == rb["res"], (pos, ra["res"], rb["res"])
        for _ in range(10):
            bad = xz.copy()
            bad[int(rng.integers(12, len(xz)))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            assert ref.xz_decode(bad, len(data) + 10)["res"] == ora.xz_decode(bad, len(data) + 10)["res"]
        # truncated: both say "need more input", not finished
        for cut in (0, 3, 11, 12, 13, 40, len(xz) // 2, len(xz) - 13, len(xz) - 12, len(xz) - 1):
            ra = ref.xz_decode(xz[:cut], len(data) + 10)
            rb = ora.xz_decode(xz[:cut], len(data) + 10)
            assert (ra["res"], ra["finished"]) == (rb["res"], rb["finished"]), cut


def test_xz_matches_system_decoder(ref, ora, corpus):
    import lzma
    data = add_x86_callsites(corpus[:300000], seed=3)
    xz = _make_xz(ref, data, 100000, 4, 4, 0)
    assert lzma.decompress(bytes(xz)) == data.tobytes()


def test_entropy_sweep(ref, ora):
    for t in (0.0, 0.3, 0.7, 1.0):
        plain = make_corpus(120000, seed=int(t * 10), mix=entropy_sweep_mix(t))
        packed, props = ref.lzma_encode(plain, level=5, dict_size=1 << 16)
        same(ref.lzma_decode(packed, props, len(plain), 0), ora.lzma_decode(packed, props, len(plain), 0))
