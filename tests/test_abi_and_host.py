"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/lzb200.h declares, keeps the reference's public struct layouts, refuses to decode without a
device (no CPU fallback), and its HOST-side front ends (property parsing, LZMA2 unit walk, xz index and
varints) agree with the oracle.  No compute call is made here."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import lzb200
from lzb200 import lib as L
from lzb200 import records as R


class CLzmaProps(C.Structure):
    _fields_ = [("lc", C.c_uint), ("lp", C.c_uint), ("pb", C.c_uint), ("dicSize", C.c_uint32)]


class CLzmaDec(C.Structure):          # reference LzmaDec.h:50-69
    _fields_ = [("prop", CLzmaProps), ("probs", C.c_void_p), ("dic", C.c_void_p), ("buf", C.c_void_p),
                ("range", C.c_uint32), ("code", C.c_uint32), ("dicPos", C.c_size_t), ("dicBufSize", C.c_size_t),
                ("processedPos", C.c_uint32), ("checkDicSize", C.c_uint32), ("state", C.c_uint),
                ("reps", C.c_uint32 * 4), ("remainLen", C.c_uint), ("needFlush", C.c_int), ("needInitState", C.c_int),
                ("numProbs", C.c_uint32), ("tempBufSize", C.c_uint), ("tempBuf", C.c_uint8 * 20)]


class CLzma2Dec(C.Structure):         # reference Lzma2Dec.h:15-25
    _fields_ = [("decoder", CLzmaDec), ("packSize", C.c_uint32), ("unpackSize", C.c_uint32), ("state", C.c_int),
                ("control", C.c_uint8), ("needInitDic", C.c_int), ("needInitState", C.c_int), ("needInitProp", C.c_int)]


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    names = [n for n in L.declared_symbols() if n != "void"]
    assert len(names) >= 45
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_public_struct_layouts_match_reference_abi():
    # SURVEY.md §8b (measured on the reference): sizeof 136 / 168 / 1792 and the field offsets
    assert C.sizeof(CLzmaDec) == 136 and C.sizeof(CLzma2Dec) == 168
    offs = {f: getattr(CLzmaDec, f).offset for f, _ in CLzmaDec._fields_}
    assert offs == dict(prop=0, probs=16, dic=24, buf=32, range=40, code=44, dicPos=48, dicBufSize=56,
                        processedPos=64, checkDicSize=68, state=72, reps=76, remainLen=92, needFlush=96,
                        needInitState=100, numProbs=104, tempBufSize=108, tempBuf=112)
    assert CLzma2Dec.packSize.offset == 136 and CLzma2Dec.needInitProp.offset == 160
    # the xz objects: compile the public header as C and ask the compiler
    import subprocess, tempfile
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "lzb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(CXzUnpacker), sizeof(CMixCoder), sizeof(CXzBlock),
         sizeof(CXzCheck), sizeof(IStateCoder), sizeof(CSha256), offsetof(CXzUnpacker, numBlocks),
         offsetof(CXzUnpacker, decoder), offsetof(CXzUnpacker, buf));
  return 0; }
"""
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "abi.c"), os.path.join(td, "abi")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.dirname(L.HEADER), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    assert got == [1792, 272, 152, 120, 40, 104, 40, 88, 768], got       # = the reference build (gcc, x86-64)
    assert R.UNIT_DTYPE.itemsize == 48 and R.RESULT_DTYPE.itemsize == 24


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.LzbError) as e:
        L.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    blobs, cs = cases.golden()
    c = cases.cases_of("lzma")[0]
    with pytest.raises(L.LzbError):
        lzb200.sdk.lzma_decode(blobs[c["blob"]], bytes.fromhex(c["props"]), c["cap"], 0)


def test_props_decode():
    lib = L.load()
    p = CLzmaProps()
    for d in range(0, 256, 7):
        raw = bytes([d, 0x34, 0x12, 0, 0])
        r = lib.LzmaProps_Decode(C.byref(p), raw, 5)
        if d >= 225:
            assert r == R.SZ_ERROR_UNSUPPORTED
        else:
            assert r == 0 and (p.lc, p.lp, p.pb) == (d % 9, (d // 9) % 5, d // 45) and p.dicSize == 4660
            assert R.parse_lzma_props(raw) == (p.lc, p.lp, p.pb, 4660)
    assert lib.LzmaProps_Decode(C.byref(p), bytes([0x5D, 0, 0, 0, 0]), 5) == 0 and p.dicSize == 4096
    assert lib.LzmaProps_Decode(C.byref(p), bytes([0x5D, 0, 0, 0]), 4) == R.SZ_ERROR_UNSUPPORTED


def test_varint_and_check_size():
    lib = L.load()
    lib.Xz_ReadVarInt.restype = C.c_uint
    lib.Xz_ReadVarInt.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
    lib.XzFlags_GetCheckSize.restype = C.c_uint
    v = C.c_uint64()
    assert lib.Xz_ReadVarInt(bytes([0x7F]), 1, C.byref(v)) == 1 and v.value == 127
    assert lib.Xz_ReadVarInt(bytes([0x80, 0x01]), 2, C.byref(v)) == 2 and v.value == 128
    assert lib.Xz_ReadVarInt(bytes([0x80, 0x00]), 2, C.byref(v)) == 0          # non-minimal encoding
    assert lib.Xz_ReadVarInt(bytes([0x80]), 1, C.byref(v)) == 0                # truncated
    assert [lib.XzFlags_GetCheckSize(t) for t in (0, 1, 4, 10)] == [0, 4, 8, 32]


def test_crc_tables_are_generated_like_the_reference():
    lib = L.load()
    lib.CrcGenerateTable()
    lib.Crc64GenerateTable()
    t32 = (C.c_uint32 * 2048).in_dll(lib, "g_CrcTable")
    t64 = (C.c_uint64 * 256).in_dll(lib, "g_Crc64Table")
    assert t32[1] == 0x77073096 and t32[255] == 0x2D02EF8D
    assert t64[1] == 0xB32E4CBE03A75F6F


def test_lzma2_unit_walk_matches_oracle(ora):
    blobs, _ = cases.golden()
    for name in ("lzma2_0", "lzma2_1", "lzma2_3", "lzma2_multi"):
        blob = blobs[name]
        prop = [c["prop"] for c in cases.cases_of("lzma2") if c["blob"] == name][0]
        mine = lzb200.sdk.lzma2_index_units(blob, prop)
        want = ora.lzma2_index(blob)
        assert mine["res"] == 0 and mine["n"] == want["n"] and mine["has_eos"] == want["has_eos"]
        assert mine["end_off"] == want["end_off"]
        for u, w in zip(mine["units"], want["units"]):
            assert (u["src_off"], u["dst_cap"]) == (w["src_off"], w["unpack_len"])
            assert u["src_len"] == w["src_len"] and u["lc"] == w["max_lclp"] and u["codec"] == R.CODEC_LZMA2
        assert np.array_equal(np.cumsum(mine["units"]["dst_cap"])[:-1], mine["units"]["dst_off"][1:])
    assert lzb200.sdk.lzma2_index_units(blobs["lzma2_multi"], 41)["res"] == R.SZ_ERROR_UNSUPPORTED
    assert lzb200.sdk.lzma2_index_units(blobs["lzma2_multi"], 20)["n"] == 4


def test_xz_backward_index_matches_oracle(ora):
    lib = L.load()
    blobs, _ = cases.golden()

    class Blk(C.Structure):      # include/lzb200.h: LzbXzBlock
        _fields_ = [("hdr_off", C.c_uint64), ("hdr_size", C.c_uint32), ("data_off", C.c_uint64),
                    ("pack_size", C.c_uint64), ("unpack_size", C.c_uint64), ("check_off", C.c_uint64),
                    ("expect", C.c_uint64), ("filter", C.c_int32), ("filter_param", C.c_uint32),
                    ("lzma2_prop", C.c_uint8), ("check_type", C.c_uint8), ("complete", C.c_uint8), ("pad_", C.c_uint8),
                    ("first_unit", C.c_uint32), ("n_units", C.c_uint32)]
    for name in ("xz0", "xz1", "xz2", "xz3"):
        xz = blobs[name]
        want = ora.xz_index(xz)
        arr = (Blk * 64)()
        n, ct = C.c_size_t(0), C.c_int(0)
        r = lib.Xz_ReadIndexBackward(xz.ctypes.data, len(xz), arr, 64, C.byref(n), C.byref(ct))
        assert r == want["res"] == 0 and n.value == len(want["blocks"]) and ct.value == want["check_type"]
        for b, w in zip(arr[: n.value], want["blocks"]):
            assert (b.hdr_off, b.hdr_size, b.data_off, b.pack_size, b.unpack_size, b.check_off) == \
                   (w["hdr_off"], w["hdr_size"], w["data_off"], w["pack_size"], w["unpack_size"], w["check_off"])
            assert (b.filter, b.filter_param, b.lzma2_prop) == (w["filter_id"], w["filter_param"], w["lzma2_prop"])
        bad = xz.copy()
        bad[-5] ^= 1
        arr2 = (Blk * 64)()
        assert lib.Xz_ReadIndexBackward(bad.ctypes.data, len(bad), arr2, 64, C.byref(n), C.byref(ct)) == ora.xz_index(bad)["res"] != 0


class LzbGeometry(C.Structure):       # csrc/lzb_kernels.h
    _fields_ = [("group", C.c_uint32), ("warps_per_cta", C.c_uint32), ("units_per_cta", C.c_uint32), ("ctas", C.c_uint32),
                ("region_bytes", C.c_uint32), ("smem_stride", C.c_uint32), ("slab_stride", C.c_uint32),
                ("shared_model", C.c_int), ("ahead", C.c_int), ("slab_bytes", C.c_size_t)]


def test_few_units_launch_plan():
    """Launches with fewer units than warps (csrc/lzb_kernels.cu, profiles/r01_j_few_units.md): one unit per warp,
    spread over all SMs, matched-literal trees in shared memory when they fit.  Pure arithmetic, no device."""
    lib = L.load()
    plan = lib.lzb_plan_few_units
    plan.restype = None
    plan.argtypes = [C.POINTER(LzbGeometry), C.c_uint32, C.c_size_t, C.c_int, C.c_int]
    optin = 232448                                  # B200: 227 KiB per CTA

    def default(region=6448, slab=8192, units=36):  # lc3/lp0/pb2: the throughput geometry of lzb_launch_geometry
        return LzbGeometry(16, units // 2, units, 148, region, region, slab, 1, 0, slab * units * 148)

    def run(n, regroup=1, dense=0, **kw):
        g = default(**kw)
        plan(C.byref(g), n, optin, regroup, dense)
        assert g.warps_per_cta * (32 // g.group) == g.units_per_cta
        assert g.units_per_cta * g.smem_stride <= optin                   # the launch fits in shared memory
        assert g.smem_stride in (g.region_bytes, g.region_bytes + g.slab_stride)
        if g.group == 32:
            assert g.ctas * g.units_per_cta >= n or g.warps_per_cta == 16 # one wave, or 16 persistent warps per SM
            assert g.units_per_cta * g.ctas * g.slab_stride <= g.slab_bytes   # inside the slab sized for the default
        return g

    g = run(65536)                                  # the bench batch: throughput mode untouched
    assert (g.group, g.warps_per_cta, g.units_per_cta, g.smem_stride, g.ahead) == (16, 18, 36, 6448, 0)
    g = run(148 * 16 + 1)                           # a little more than one wave of warps: still one unit per warp, persistent
    assert (g.group, g.warps_per_cta, g.ahead) == (32, 16, 1)
    g = run(148 * 29)
    assert (g.group, g.warps_per_cta) == (32, 16)
    g = run(148 * 29 + 1)                           # from here on two units per warp
    assert (g.group, g.warps_per_cta, g.ahead) == (16, 18, 0)
    g = run(2220)                                   # xz4m: 15 per SM, trees fit (15 x 14,640 B); one unit per warp: look-ahead decoder
    assert (g.group, g.warps_per_cta, g.smem_stride, g.ahead) == (32, 15, 6448 + 8192, 1)
    g = run(148 * 16)                               # 16 per SM: one warp each, trees stay in the slab
    assert (g.group, g.warps_per_cta, g.smem_stride) == (32, 16, 6448)
    g = run(1024)                                   # spread: 7 warps on every SM instead of 16 on 64; look-ahead decoder
    assert (g.group, g.warps_per_cta, g.smem_stride, g.ahead) == (32, 7, 14640, 1)
    for n in (1, 148):
        g = run(n)
        assert (g.group, g.warps_per_cta, g.smem_stride, g.ahead) == (32, 1, 14640, 1)
    g = run(1024, dense=1)                          # LZB_SMEM_MATCHED=0: the old packing
    assert (g.group, g.warps_per_cta, g.smem_stride) == (32, 16, 6448)
    g = run(1024, regroup=0)                        # LZB_GROUP forced: lanes per unit are the caller's
    assert (g.group, g.warps_per_cta, g.smem_stride) == (16, 18, 6448)
    # LZMA2 maxima (lc+lp = 4, pb = 4): 11,888 B region, 16 KiB of matched trees -> 8 units per SM with the trees
    g = run(148 * 8, region=11888, slab=16384, units=18)
    assert (g.group, g.warps_per_cta, g.smem_stride) == (32, 8, 11888 + 16384)
    g = run(148 * 9, region=11888, slab=16384, units=18)
    assert (g.group, g.warps_per_cta, g.smem_stride) == (32, 9, 11888)


def test_lzma2_unit_walk_fuzz_against_oracle(ora):
    """Byte mutations of the golden LZMA2 streams (aimed at the chunk headers half of the time): the host walker
    accepts / rejects exactly the streams the oracle's walker does and cuts the same units.  When the last chunk
    runs past the buffer the two helpers define `end_off` differently (end of the last complete chunk here, the
    buffer length there) -- the decode results for such streams are pinned by the golden 'cut' cases on the GPU."""
    blobs, _ = cases.golden()
    rng = np.random.default_rng(11)
    rejected = 0
    for name in ("lzma2_0", "lzma2_1", "lzma2_3", "lzma2_multi"):
        blob = blobs[name]
        prop = [c["prop"] for c in cases.cases_of("lzma2") if c["blob"] == name][0]
        for _ in range(150):
            b = blob.copy()
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, len(b))) if rng.random() < 0.4 else int(rng.integers(0, 8))
                b[pos] = rng.integers(0, 256)
            mine = lzb200.sdk.lzma2_index_units(b, prop)
            want = ora.lzma2_index(b)
            if want["n"] == -1:
                assert mine["res"] != 0
                rejected += 1
                continue
            assert mine["res"] == 0 and (mine["n"], mine["has_eos"]) == (want["n"], want["has_eos"])
            complete = want["has_eos"] or mine["end_off"] == want["end_off"]
            if complete:
                assert mine["end_off"] == want["end_off"]
                for u, w in zip(mine["units"], want["units"]):
                    assert (u["src_off"], u["dst_cap"], u["src_len"], u["lc"]) == \
                           (w["src_off"], w["unpack_len"], w["src_len"], w["max_lclp"])
    assert rejected > 50


def test_xz_backward_index_fuzz_against_oracle(ora):
    """Mutations aimed at the stream footer, the index and the block headers of the golden containers: the backward
    index reader returns the oracle's result code and, when it accepts, the same block table."""
    lib = L.load()
    blobs, _ = cases.golden()

    class Blk(C.Structure):      # include/lzb200.h: LzbXzBlock
        _fields_ = [("hdr_off", C.c_uint64), ("hdr_size", C.c_uint32), ("data_off", C.c_uint64),
                    ("pack_size", C.c_uint64), ("unpack_size", C.c_uint64), ("check_off", C.c_uint64),
                    ("expect", C.c_uint64), ("filter", C.c_int32), ("filter_param", C.c_uint32),
                    ("lzma2_prop", C.c_uint8), ("check_type", C.c_uint8), ("complete", C.c_uint8), ("pad_", C.c_uint8),
                    ("first_unit", C.c_uint32), ("n_units", C.c_uint32)]
    rng = np.random.default_rng(5)
    rejected = accepted = 0
    for name in ("xz0", "xz1", "xz2", "xz3"):
        xz = blobs[name]
        hdrs = [b["hdr_off"] for b in ora.xz_index(xz)["blocks"]]
        for _ in range(120):
            b = xz.copy()
            for _ in range(int(rng.integers(1, 3))):
                r = rng.random()
                if r < 0.5:
                    pos = len(b) - 1 - int(rng.integers(0, 80))
                elif r < 0.8:
                    pos = int(hdrs[int(rng.integers(0, len(hdrs)))]) + int(rng.integers(0, 24))
                elif r < 0.9:
                    pos = int(rng.integers(0, 12))
                else:
                    pos = int(rng.integers(0, len(b)))
                b[pos] = rng.integers(0, 256)
            want = ora.xz_index(b)
            arr = (Blk * 64)()
            n, ct = C.c_size_t(0), C.c_int(0)
            r = lib.Xz_ReadIndexBackward(b.ctypes.data, len(b), arr, 64, C.byref(n), C.byref(ct))
            assert r == want["res"]
            if r != 0:
                rejected += 1
                continue
            accepted += 1
            assert n.value == len(want["blocks"]) and ct.value == want["check_type"]
            for bb, w in zip(arr[: n.value], want["blocks"]):
                assert (bb.hdr_off, bb.hdr_size, bb.data_off, bb.pack_size, bb.unpack_size, bb.check_off) == \
                       (w["hdr_off"], w["hdr_size"], w["data_off"], w["pack_size"], w["unpack_size"], w["check_off"])
                assert (bb.filter, bb.filter_param, bb.lzma2_prop) == (w["filter_id"], w["filter_param"], w["lzma2_prop"])
    assert rejected > 100 and accepted > 50
