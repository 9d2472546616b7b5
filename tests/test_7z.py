"""The 7z folder path (BASELINE config 5): SzArchive_Open / SzArchive_ExtractAll must answer
what the reference's SzArEx_Open and SzArEx_Extract answer -- for the database (file table, names, folder
map), for every file's bytes, and for the result codes of damaged archives.

Golden: tests/golden/sz.json (tools/make_golden_7z.py): archives written by tools/sz_writer.py from payloads
the reference's encoders made (Copy / LZMA / LZMA2, BCJ x86, ARM, BCJ2 chains, Delta-as-unsupported, packed header, many
small folders from low entropy to random, empty files and directories), pristine and with 29 mutations each.

CPU part : vectors vs the live reference (needref); the writer's own listing; the public structs.
GPU part : liblzb200.so vs the vectors.
"""
import json
import os
import zlib

import numpy as np
import pytest

from tools.make_golden_7z import mutate, summarize

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    j = json.load(open(os.path.join(HERE, "golden", "sz.json")))
    z = np.load(os.path.join(HERE, "golden", "sz.npz"))
    return j["cases"], {k: z[k].tobytes() for k in z.files}


def compare(got, want, tag):
    assert got["open_res"] == want["open_res"], (tag, got["open_res"], want["open_res"])
    if want["open_res"] != 0:
        return
    assert (got["num_files"], got["num_folders"]) == (want["num_files"], want["num_folders"]), tag
    for k, (a, b) in enumerate(zip(got["files"], want["files"])):
        assert a == b, (tag, k, a, b)


@pytest.mark.needref
def test_golden_equals_live_reference(ref):
    cases, blobs = golden()
    for c in cases[::4]:
        m = tuple(c["mut"]) if c["mut"] else None
        compare(summarize(ref.sz_extract_all(mutate(blobs[c["blob"]], m))), c, (c["blob"], c["mut"]))


def test_golden_is_meaningful():
    cases, blobs = golden()
    assert len(cases) >= 150
    assert {c["open_res"] for c in cases} >= {0, 3, 4, 6, 16, 17}
    assert {f["res"] for c in cases for f in c["files"]} >= {0, 1, 3, 4}
    whole = [c for c in cases if c["mut"] is None]
    assert len(whole) == 6 and sum(c["num_folders"] for c in whole) > 70
    assert any(f["is_dir"] for c in whole for f in c["files"]) and any(not f["has_stream"] and not f["is_dir"] for c in whole for f in c["files"])


def test_public_structs_have_the_documented_size():
    import lzb200
    assert lzb200.sdk.SZ_FOLDER_DTYPE.itemsize == 64 and lzb200.sdk.SZ_FILE_DTYPE.itemsize == 56


@pytest.mark.gpu
@pytest.mark.parametrize("blob", ["a_plain", "a_packed_header", "b_many", "c_delta_unsupported", "d_empty_only", "e_bcj2"])
def test_product_matches_golden(blob):
    import lzb200
    cases, blobs = golden()
    mine = [c for c in cases if c["blob"] == blob]
    assert mine
    for c in mine:
        m = tuple(c["mut"]) if c["mut"] else None
        d = lzb200.sdk.sz_extract_all(mutate(blobs[blob], m))
        compare(summarize(d), c, (blob, c["mut"]))


@pytest.mark.gpu
def test_all_folders_in_one_pass_and_bytes_match_the_writer():
    """fresh archive (not in the golden set): 200 folders, every byte compared with what was stored"""
    import lzb200
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("needs the reference encoder to write a fresh archive")
    from tools import sz_writer as W
    ref = refbind.ref()
    rng = np.random.default_rng(1)
    folders = []
    for i in range(200):
        n = int(rng.integers(100, 20000))
        data = np.where(rng.random(n) < i / 199.0, rng.integers(0, 256, n), (np.arange(n) // 5) % 11).astype(np.uint8).tobytes()
        folders.append(W.Folder("lzma2" if i & 1 else "lzma", [(f"n{i}.bin", data)], filt="bcj" if i % 7 == 0 else None, dict_size=1 << 15))
    arc = W.write_archive(ref, folders, encode_header=True)
    import os
    # pageable buffers (staged passes), then page-locked ones with the streamed decode pass forced on
    for pinned in (False, True):
        os.environ["LZB_FLOW"] = "1"
        try:
            d = lzb200.sdk.sz_extract_all(arc, pinned=pinned)
        finally:
            os.environ.pop("LZB_FLOW", None)
        assert d["open_res"] == 0 and d["num_folders"] == 200 and all(r == 0 for r in d["folder_res"])
        for f, (nm, data) in zip(d["files"], W.expected_listing(folders)):
            assert f["res"] == 0 and f["name"] == nm and d["out"][f["off"]:f["off"] + f["size"]].tobytes() == data
    # a damaged member must still be found by the check pass that follows the streamed launch
    bad = bytearray(arc)
    bad[len(arc) // 2] ^= 0x20
    os.environ["LZB_FLOW"] = "1"
    try:
        a, b = lzb200.sdk.sz_extract_all(bytes(bad), pinned=False), lzb200.sdk.sz_extract_all(bytes(bad), pinned=True)
    finally:
        os.environ.pop("LZB_FLOW", None)
    assert a["open_res"] == b["open_res"] and [f["res"] for f in a["files"]] == [f["res"] for f in b["files"]]
    assert a["open_res"] != 0 or any(f["res"] != 0 for f in a["files"])


# ---------------------------------------------------------------------------------------- hostile headers (CPU)
def _plain_header_archive():
    """Three stored folders with six files, plain (not packed) header: SzArchive_Open needs no GPU for it."""
    from tools import sz_writer as W
    rng = np.random.default_rng(9)
    folders = []
    for k in range(3):
        files = [(f"d{k}/f{j}.bin", rng.integers(0, 256, 300 + 50 * j, dtype=np.uint8).tobytes()) for j in range(2)]
        f = W.Folder("copy", files)
        f.plain = b"".join(d for _, d in files)
        f.packed = f.plain
        folders.append(f)
    return W.write_archive(None, folders, empty_files=("e.txt",), dirs=("dir",))


def _reseal(arc: bytearray):
    """Recompute the two CRCs of the signature header so that a mutated database is actually parsed."""
    import struct
    next_off, next_size = struct.unpack_from("<QQ", arc, 12)
    lo = 32 + next_off
    if lo + next_size <= len(arc):
        struct.pack_into("<I", arc, 28, zlib.crc32(bytes(arc[lo:lo + next_size])))
    struct.pack_into("<I", arc, 8, zlib.crc32(bytes(arc[12:32])))


def test_hostile_databases_are_refused_on_the_host():
    """Sizes, counts and offsets in a 7z database are the archive's word.  Whatever they say, SzArchive_Open answers
    with a result code (no exception through the C ABI, no allocation of the archive's choosing), and a database it
    accepts places every folder and every file inside SzArchive_UnpackSize -- nothing out of range can reach the GPU."""
    import ctypes as C
    import struct
    import lzb200
    from lzb200.sdk import SZ_FILE_DTYPE, SZ_FOLDER_DTYPE
    L = lzb200.lib.load()
    L.SzArchive_Open.restype = C.c_int
    L.SzArchive_Open.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.SzArchive_Free.argtypes = [C.c_void_p]
    for f in ("SzArchive_NumFolders", "SzArchive_NumFiles"):
        getattr(L, f).restype = C.c_size_t
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("SzArchive_Folders", "SzArchive_Files"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.SzArchive_UnpackSize.restype = C.c_uint64
    L.SzArchive_UnpackSize.argtypes = [C.c_void_p]

    def open_and_check(arc: bytes):
        h = C.c_void_p(None)
        r = L.SzArchive_Open(None, arc, len(arc), C.byref(h))
        if r != 0:
            assert not h.value
            return r
        try:
            total = int(L.SzArchive_UnpackSize(h))
            nfo, nfi = L.SzArchive_NumFolders(h), L.SzArchive_NumFiles(h)
            if nfo:
                fo = np.ctypeslib.as_array(C.cast(L.SzArchive_Folders(h), C.POINTER(C.c_uint8)), shape=(nfo * 64,)).view(SZ_FOLDER_DTYPE)
                assert all(int(x["out_off"]) + int(x["unpack_size"]) <= total for x in fo)
            if nfi:
                fi = np.ctypeslib.as_array(C.cast(L.SzArchive_Files(h), C.POINTER(C.c_uint8)), shape=(nfi * 56,)).view(SZ_FILE_DTYPE)
                assert nfi < 1 << 20
                assert all(int(x["folder"]) == 0xFFFFFFFF or int(x["folder"]) < nfo for x in fi)
        finally:
            L.SzArchive_Free(h)
        return 0

    good = _plain_header_archive()
    assert open_and_check(good) == 0
    next_off, next_size = struct.unpack_from("<QQ", good, 12)
    lo = 32 + next_off
    rng = np.random.default_rng(17)
    seen = set()
    for trial in range(3000):
        arc = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            at = lo + int(rng.integers(0, next_size))
            arc[at] = int(rng.choice([0, 1, 0x7F, 0x80, 0xC0, 0xE0, 0xF0, 0xFF, int(rng.integers(0, 256))]))
        _reseal(arc)
        seen.add(open_and_check(bytes(arc)))
    assert 0 in seen and len(seen) >= 3          # the mutations do reach the parser's error paths
    # counts of the archive's choosing: 2^31 - 1 files / folders / pack streams in a header of a few bytes
    from tools import sz_writer as W
    huge = W.num(0x7FFFFFFF)
    for hdr in (bytes([W.K_HEADER, W.K_FILES]) + huge + bytes([W.K_END, W.K_END]),
                bytes([W.K_HEADER, W.K_MAIN_STREAMS, 6]) + W.num(0) + huge + bytes([9]),
                bytes([W.K_HEADER, W.K_MAIN_STREAMS, 7, 11]) + huge + bytes([0])):
        start = struct.pack("<QQI", 0, len(hdr), zlib.crc32(hdr))
        arc = W.SIG + bytes([0, 3]) + struct.pack("<I", zlib.crc32(start)) + start + hdr
        assert open_and_check(arc) != 0
