"""CPU suite over the committed golden vectors (no GPU, no /root/reference needed):

 * the oracle (oracle/*.c) reproduces every vector the reference produced         -> oracle stays pinned
 * the device decode logic, compiled for the host with -DLZB_EMU (tests/emu, a TEST-ONLY single-lane
   emulation of csrc/lzb_core.cuh + lzb_post.cuh), reproduces them too            -> kernel logic checked
   before any GPU time is spent.
"""
import numpy as np
import pytest

import cases
import emubind


@pytest.fixture(scope="module", params=["lean", "ahead"])
def emu(request):
    """Both forms of the decoder the kernels instantiate: lean (many units per SM) and look-ahead (a warp per unit)."""
    emubind.lib()
    emubind.AHEAD = 1 if request.param == "ahead" else 0
    yield emubind.Emu()
    emubind.AHEAD = 0


@pytest.mark.parametrize("kind", ["lzma", "lzma2"])
def test_oracle_reproduces_golden_decodes(ora, kind):
    blobs, _ = cases.golden()
    cs = cases.cases_of(kind)
    assert len(cs) > 50
    for c in cs:
        cases.check_decode_case(ora, c, blobs)


def test_oracle_reproduces_golden_xz(ora):
    blobs, _ = cases.golden()
    for c in cases.cases_of("xz"):
        cases.check_xz_case(ora, c, blobs)


@pytest.mark.parametrize("kind", ["x86", "delta", "crc"])
def test_oracle_reproduces_golden_filters(ora, kind):
    blobs, _ = cases.golden()
    for c in cases.cases_of(kind):
        cases.check_filter_case(ora, c, blobs)


@pytest.mark.parametrize("kind", ["lzma", "lzma2"])
def test_kernel_logic_reproduces_golden_decodes(emu, kind):
    blobs, _ = cases.golden()
    for c in cases.cases_of(kind):
        cases.check_decode_case(emu, c, blobs)


@pytest.mark.parametrize("kind", ["x86", "delta", "crc"])
def test_kernel_logic_reproduces_golden_filters(emu, kind):
    blobs, _ = cases.golden()
    for c in cases.cases_of(kind):
        cases.check_filter_case(emu, c, blobs)


def test_branch_converters_and_delta_encode(ora, emu):
    """ARM / ARMT / PPC / SPARC / IA64 and Delta_Encode: oracle and kernel logic vs the reference's outputs."""
    cs, blobs = cases.bra_golden()
    assert len(cs) > 300
    for impl in (ora, emu):
        for c in cs:
            cases.check_bra_case(impl, c, blobs)


def test_x86_converters_on_candidate_dense_buffers(emu, ora):
    """Both x86 BCJ converters of lzb_post.cuh (per-lane segments under the lean form, the tile scan under the look-ahead
    form) against the oracle on buffers that are mostly opcode candidates and fill bytes: every carried state, both
    directions, sizes around the 16-byte tile, unaligned buffer starts (the emulation's arenas start at offset 5)."""
    rng = np.random.default_rng(20260)
    for it in range(1500):
        n = int(rng.choice([5, 6, 9, 15, 16, 17, 20, 31, 32, 33, 47, 48, 49, 64, 100, 257, 1000]))
        if it % 3 == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif it % 3 == 1:
            d = rng.choice(np.array([0xE8, 0xE9, 0x00, 0xFF, 0x12, 0x80], dtype=np.uint8), n, p=[.25, .15, .25, .2, .1, .05])
        else:
            d = rng.choice(np.array([0xE8, 0x00, 0xFF], dtype=np.uint8), n)
        ip, st, enc = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 8)), int(rng.integers(0, 2))
        w, g = ora.x86_convert(d, ip, st, enc), emu.x86_convert(d, ip, st, enc)
        assert np.array_equal(w[0], g[0]) and w[1:] == g[1:], (n, ip, st, enc)


@pytest.mark.needref
def test_branch_converter_golden_equals_live_reference(ref):
    cs, blobs = cases.bra_golden()
    for c in cs[::5]:
        cases.check_bra_case(ref, c, blobs)


def test_kernel_logic_fused_filter_and_check(emu, ora):
    """decode -> BCJ -> CRC64 in one pass (what an xz block with one unit asks of a warp)."""
    import lzb200
    R = lzb200.records
    blobs, cs = cases.golden()
    c = [x for x in cs if x["kind"] == "xz" and x["blob"] == "xz0" and x["mut"] is None][0]
    idx = ora.xz_index(blobs["xz0"])
    assert idx["res"] == 0 and idx["check_type"] == 4
    xz = blobs["xz0"]
    total = sum(b["unpack_size"] for b in idx["blocks"])
    out = np.zeros(total + 64, np.uint8)
    units = R.make_units(len(idx["blocks"]))
    off = 0
    for i, b in enumerate(idx["blocks"]):
        u = units[i]
        u["src_off"], u["src_len"] = b["data_off"], b["pack_size"]
        u["dst_off"], u["dst_cap"] = off, b["unpack_size"]
        u["dict_size"] = (2 | (b["lzma2_prop"] & 1)) << (b["lzma2_prop"] // 2 + 11)
        u["codec"], u["finish"] = R.CODEC_LZMA2, R.LZMA_FINISH_END
        u["filter"], u["filter_param"], u["check"], u["flags"] = b["filter_id"], b["filter_param"], 4, R.UNIT_VERIFY
        u["expect"] = int.from_bytes(xz[b["check_off"]:b["check_off"] + 8].tobytes(), "little")
        off += b["unpack_size"]
    res = emubind.decode_units(np.concatenate((xz, np.zeros(16, np.uint8))), out, units)
    assert (res["res"] == 0).all() and (res["status"] == R.LZMA_STATUS_FINISHED_WITH_MARK).all()
    import zlib
    assert zlib.crc32(out[:total].tobytes()) == c["plain_crc"]
    # a wrong stored digest must surface as SZ_ERROR_CRC for that unit only
    units["expect"][1] ^= 1
    res = emubind.decode_units(np.concatenate((xz, np.zeros(16, np.uint8))), out, units)
    assert res["res"].tolist() == [0, R.SZ_ERROR_CRC] + [0] * (len(units) - 2)


@pytest.mark.needref
def test_fused_x86_filter_starts_a_block_without_history(emu, ref):
    """A BCJ filter fused behind a decode starts every block with history 0 -- also when the unit's `expect` field
    (the stored digest there, the incoming history only for stand-alone x86_Convert calls) has bits 32..34 set.
    Blocks that begin with E8 candidates at offsets 0..2 are the ones that can tell."""
    import lzb200
    R = lzb200.records
    rng = np.random.default_rng(3)
    done = 0
    for trial in range(64):
        plain = rng.integers(0, 256, 3000, dtype=np.uint8)
        for at in (0, 1, 2, 7, 9):
            plain[at] = 0xE8
            plain[at + 4] = 0x00 if (at + trial) & 1 else 0xFF
        digest = ref.crc64(plain)
        if ((digest >> 32) & 7) == 0:
            continue
        pk, prop = ref.lzma2_encode(ref.x86_filter_whole(plain, 0, 1), level=1, dict_size=1 << 16)
        u = R.make_units(1)
        u["src_len"], u["dst_cap"], u["dict_size"] = len(pk), len(plain), (2 | (prop & 1)) << (prop // 2 + 11)
        u["codec"], u["finish"], u["filter"], u["check"], u["flags"], u["expect"] = R.CODEC_LZMA2, R.LZMA_FINISH_END, 4, 4, R.UNIT_VERIFY, digest
        out = np.zeros(len(plain) + 64, np.uint8)
        res = emubind.decode_units(np.concatenate((pk, np.zeros(16, np.uint8))), out, u)
        assert int(res["res"][0]) == 0 and np.array_equal(out[: len(plain)], plain), trial
        done += 1
    assert done >= 20


@pytest.mark.needref
def test_lzma86_golden_equals_live_reference(ref):
    import ctypes as C
    import json
    import os
    from oracle import refbind
    from tools.make_golden_lzma86 import bind, decode, mutate
    here = os.path.dirname(os.path.abspath(__file__))
    cs = json.load(open(os.path.join(here, "golden", "lzma86.json")))["cases"]
    z = np.load(os.path.join(here, "golden", "lzma86.npz"))
    L = bind(C.CDLL(refbind.REF_SO))
    for c in cs:
        m = tuple(c["mut"]) if c["mut"] else None
        got = decode(L, mutate(z[c["blob"]].tobytes(), m), c["cap"])
        assert got == {k: c[k] for k in got}, (c["blob"], c["mut"])


def test_second_opinion_liblzma_agrees_with_the_oracle(ora):
    """liblzma (python `lzma`) as an independent decoder of the golden raw-LZMA streams
    (.lzma framing: 5 props bytes + 8 size bytes + stream) and of the golden xz containers."""
    import lzma
    blobs, _ = cases.golden()
    n = 0
    for c in cases.cases_of("lzma"):
        if c["mut"] is not None or c["res"] != 0 or c["dest_len"] != c["cap"] or c["cap"] == 0 or c["finish"] != 0:
            continue
        if c["src_len"] != len(blobs[c["blob"]]):
            continue                                        # an end mark follows: .lzma with a known size does not allow one
        props = bytes.fromhex(c["props"])
        if props[0] // 9 // 5 > 4 or (props[0] % 9) + (props[0] // 9) % 5 > 4:
            continue                                        # liblzma refuses lc + lp > 4
        alone = props + int(c["cap"]).to_bytes(8, "little") + blobs[c["blob"]].tobytes()
        want = lzma.LZMADecompressor(format=lzma.FORMAT_ALONE).decompress(alone, max_length=c["cap"])
        got = ora.lzma_decode(blobs[c["blob"]], props, c["cap"], 0)
        assert got["res"] == 0 and got["out"][: got["dest_len"]].tobytes() == want, c["blob"]
        n += 1
    assert n >= 5
    for c in cases.cases_of("xz"):
        if c["mut"] is None:
            want = lzma.decompress(blobs[c["blob"]].tobytes(), format=lzma.FORMAT_XZ)
            got = ora.xz_decode(blobs[c["blob"]], c["cap"])
            assert got["res"] == 0 and got["out"][: got["dest_len"]].tobytes() == want, c["blob"]


def test_symbol_stats_parser_decodes_golden_streams():
    """tools/symbol_stats.py (the analysis tool behind profiles/r01_l_symbol_stats.md) is an independent plain LZMA
    symbol parser: its output must reproduce the golden streams' bytes (CRC recorded from the reference)."""
    import os
    import sys
    import zlib
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tools import symbol_stats
    blobs, cs = cases.golden()
    done = 0
    for c in cs:
        if not (c["kind"] == "lzma" and c["mut"] is None and c["finish"] == 0 and c["res"] == 0
                and 0 < c["dest_len"] == c["cap"] <= 65536):
            continue
        props = bytes.fromhex(c["props"])
        d = props[0]
        lc, lp, pb = d % 9, (d // 9) % 5, d // 45
        out, counts, _ = symbol_stats.parse(bytes(blobs[c["blob"]]), lc, lp, pb, c["cap"])
        assert zlib.crc32(out) == c["out_crc"], c["blob"]
        assert counts["model_bits"] > 0
        done += 1
    assert done >= 6
