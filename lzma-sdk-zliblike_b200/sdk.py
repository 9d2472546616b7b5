"""Host-side mirror of the reference's decode API, one thin Python function per C entry point.

Every function here calls lib/liblzb200.so through its C-ABI (include/lzb200.h) and returns the same
observable tuple the reference call produces -- ``res`` (SRes), ``status``, ``dest_len``, ``src_len`` and
the output bytes -- so the parity tests read like calls into the reference.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from . import records as R


def _u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if len(b) else np.zeros(0, dtype=np.uint8)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- raw LZMA
def lzma_decode(packed, props, dest_cap: int, finish_mode: int = R.LZMA_FINISH_ANY) -> dict:
    """LzmaDecode (reference LzmaDec.h:221)."""
    L = _lib.load()
    src, pr = _u8(packed), _u8(props)
    dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
    dl, sl, st = C.c_size_t(dest_cap), C.c_size_t(len(src)), C.c_int(0)
    r = L.LzmaDecode(_p(dst), C.byref(dl), _p(src), C.byref(sl), _p(pr), len(pr), finish_mode, C.byref(st), None)
    if r == R.SZ_ERROR_FAIL:
        raise _lib.LzbError(_lib.last_error())
    return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])


def lzma_uncompress(packed, props, dest_cap: int) -> dict:
    """LzmaUncompress (reference LzmaLib.h:128)."""
    L = _lib.load()
    src, pr = _u8(packed), _u8(props)
    dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
    dl, sl = C.c_size_t(dest_cap), C.c_size_t(len(src))
    r = L.LzmaUncompress(_p(dst), C.byref(dl), _p(src), C.byref(sl), _p(pr), len(pr))
    if r == R.SZ_ERROR_FAIL:
        raise _lib.LzbError(_lib.last_error())
    return dict(res=r, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])


def lzma_uncompress_batch(ctx, packed_list, props, dest_caps) -> list[dict]:
    """LzmaUncompressBatch: n LzmaUncompress calls in one GPU pass.  props: one 5-byte record or one per stream."""
    L = _lib.load()
    n = len(packed_list)
    srcs = [_u8(x) for x in packed_list]
    dsts = [np.zeros(max(int(c), 1), dtype=np.uint8) for c in dest_caps]
    pr = _u8(props)
    stride = 0 if len(pr) == 5 and n != 1 else (5 if len(pr) == 5 * n else 0)
    src_ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in srcs])
    dst_ptrs = (C.c_void_p * n)(*[d.ctypes.data for d in dsts])
    sl = (C.c_size_t * n)(*[len(s) for s in srcs])
    dl = (C.c_size_t * n)(*[int(c) for c in dest_caps])
    res = (C.c_int * n)()
    r = L.LzmaUncompressBatch(ctx.h if ctx else None, n, dst_ptrs, dl, src_ptrs, sl, _p(pr), stride, res)
    if r != R.SZ_OK:
        raise _lib.LzbError(f"LzmaUncompressBatch: {r}: {_lib.last_error()}")
    return [dict(res=res[i], dest_len=dl[i], src_len=sl[i], out=dsts[i][:int(dest_caps[i])]) for i in range(n)]


# ----------------------------------------------------------------------------- LZMA2
def lzma2_index_units(packed, prop: int, max_units: int = 1 << 20) -> dict:
    """Lzma2_IndexUnits: host walk of the chunk headers, cut at dictionary resets."""
    L = _lib.load()
    src = _u8(packed)
    units = np.zeros(max_units, dtype=R.UNIT_DTYPE)
    n, end, eos = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    r = L.Lzma2_IndexUnits(_p(src), len(src), prop, _p(units), max_units, C.byref(n), C.byref(end), C.byref(eos))
    return dict(res=r, n=n.value, units=units[: min(n.value, max_units)].copy(), end_off=end.value, has_eos=bool(eos.value))


def lzma2_decode(packed, prop: int, dest_cap: int, finish_mode: int = R.LZMA_FINISH_END, ctx=None) -> dict:
    """Whole-buffer Lzma2Dec_DecodeToDic contract (reference Lzma2Dec.h:51), units decoded in parallel."""
    L = _lib.load()
    src = _u8(packed)
    dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
    dl, sl, st = C.c_size_t(dest_cap), C.c_size_t(len(src)), C.c_int(0)
    r = L.Lzma2DecodeParallel(ctx.h if ctx else None, _p(dst), C.byref(dl), _p(src), C.byref(sl), prop, finish_mode,
                              C.byref(st))
    if r == R.SZ_ERROR_FAIL:
        raise _lib.LzbError(_lib.last_error())
    return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])


# ----------------------------------------------------------------------------- xz
def xz_decode(data, dest_cap: int, ctx=None, out=None) -> dict:
    """Whole-buffer XzUnpacker_Code contract (reference Xz.h:223-246).  out: optional preallocated (pinned) buffer."""
    L = _lib.load()
    src = _u8(data)
    dst = out if out is not None else np.zeros(max(dest_cap, 1), dtype=np.uint8)
    assert dst.dtype == np.uint8 and dst.size >= dest_cap
    dl, sl, st = C.c_size_t(dest_cap), C.c_size_t(len(src)), C.c_int(0)
    nb, fin = C.c_uint64(0), C.c_int(0)
    r = L.XzDecodeParallel(ctx.h if ctx else None, _p(dst), C.byref(dl), _p(src), C.byref(sl), C.byref(st),
                           C.byref(nb), C.byref(fin))
    if r == R.SZ_ERROR_FAIL:
        raise _lib.LzbError(_lib.last_error())
    return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, num_blocks=nb.value,
                finished=bool(fin.value), out=dst[:dest_cap])


# ----------------------------------------------------------------------------- checks / filters
def crc32(data) -> int:
    b = _u8(data)
    return int(_lib.load().CrcCalc(_p(b), len(b)))


def crc32_update(v: int, data) -> int:
    b = _u8(data)
    return int(_lib.load().CrcUpdate(v, _p(b), len(b)))


def crc64(data) -> int:
    b = _u8(data)
    return int(_lib.load().Crc64Calc(_p(b), len(b)))


def crc64_update(v: int, data) -> int:
    b = _u8(data)
    return int(_lib.load().Crc64Update(v, _p(b), len(b)))


def x86_convert(data, ip: int = 0, state: int = 0, encoding: int = 0):
    """x86_Convert (reference Bra.h:58): returns (bytes, processed, state)."""
    buf = _u8(data).copy()
    st = C.c_uint32(state)
    n = _lib.load().x86_Convert(_p(buf), len(buf), ip, C.byref(st), encoding)
    return buf, int(n), st.value


def delta_decode(data, dist: int, state=None):
    """Delta_Decode (reference Delta.h:17): returns (bytes, state)."""
    buf = _u8(data).copy()
    st = np.zeros(256, dtype=np.uint8) if state is None else _u8(state).copy()
    _lib.load().Delta_Decode(_p(st), dist, _p(buf), len(buf))
    return buf, st


def delta_encode(data, dist: int, state=None):
    """Delta_Encode (reference Delta.h:16): returns (bytes, state)."""
    buf = _u8(data).copy()
    st = np.zeros(256, dtype=np.uint8) if state is None else _u8(state).copy()
    _lib.load().Delta_Encode(_p(st), dist, _p(buf), len(buf))
    return buf, st


_BRA = {5: "PPC_Convert", 6: "IA64_Convert", 7: "ARM_Convert", 8: "ARMT_Convert", 9: "SPARC_Convert"}


def bra_convert(filter_id: int, data, ip: int = 0, encoding: int = 0):
    """ARM/ARMT/PPC/SPARC/IA64_Convert (reference Bra.h:59-63) by xz filter id: returns (bytes, processed)."""
    buf = _u8(data).copy()
    f = getattr(_lib.load(), _BRA[filter_id])
    f.restype = C.c_size_t
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int]
    n = f(_p(buf), len(buf), ip, encoding)
    return buf, int(n)


# ----------------------------------------------------------------------------- units (batched extension)
def decode_units_host(ctx, packed: np.ndarray, out: np.ndarray, units: np.ndarray) -> np.ndarray:
    """Lzb_DecodeUnitsHost over numpy arenas (pinned or pageable)."""
    L = _lib.load()
    assert units.dtype == R.UNIT_DTYPE and packed.dtype == np.uint8 and out.dtype == np.uint8
    res = np.zeros(len(units), dtype=R.RESULT_DTYPE)
    r = L.Lzb_DecodeUnitsHost(ctx.h, _p(packed), packed.size, _p(out), out.size, _p(units), _p(res), len(units))
    if r != R.SZ_OK:
        raise _lib.LzbError(f"Lzb_DecodeUnitsHost: {r}: {_lib.last_error()}")
    return res


def decode_units_device(ctx, d_packed_ptr: int, d_out_ptr: int, d_units_ptr: int, d_results_ptr: int, n: int,
                        max_lclp: int, max_pb: int, stream_ptr: int = 0) -> None:
    """Lzb_DecodeUnitsDevice: arenas already resident in HBM; asynchronous on the given cudaStream_t."""
    L = _lib.load()
    r = L.Lzb_DecodeUnitsDevice(ctx.h, d_packed_ptr, d_out_ptr, d_units_ptr, d_results_ptr, n, max_lclp, max_pb,
                                stream_ptr)
    if r != R.SZ_OK:
        raise _lib.LzbError(f"Lzb_DecodeUnitsDevice: {r}: {_lib.last_error()}")


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8 array backed by Lzb_HostAlloc (page-locked).  Freed when the array is garbage collected."""
    L = _lib.load()
    p = L.Lzb_HostAlloc(nbytes)
    if not p:
        raise _lib.LzbError(_lib.last_error())
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
    _PINNED.append((p, buf))
    return np.frombuffer(buf, dtype=np.uint8, count=nbytes)


_PINNED: list = []


def pinned_free_all() -> None:
    """Release every buffer handed out by pinned_empty (arrays made from them must no longer be used)."""
    L = _lib.load()
    while _PINNED:
        p, _ = _PINNED.pop()
        L.Lzb_HostFree(p)


def raw_lzma_units(src_offs, src_lens, dst_offs, dst_caps, props, finish: int = R.LZMA_FINISH_ANY) -> np.ndarray:
    """Unit table for a batch of raw LZMA streams that share one 5-byte property record."""
    lc, lp, pb, ds = R.parse_lzma_props(props)
    n = len(src_lens)
    u = R.make_units(n)
    u["src_off"], u["src_len"], u["dst_off"], u["dst_cap"] = src_offs, src_lens, dst_offs, dst_caps
    u["dict_size"], u["codec"], u["lc"], u["lp"], u["pb"], u["finish"] = ds, R.CODEC_LZMA, lc, lp, pb, finish
    return u


# ----------------------------------------------------------------------------- 7z folder path
SZ_FOLDER_DTYPE = np.dtype([("pack_off", "<u8"), ("pack_size", "<u8"), ("unpack_size", "<u8"), ("out_off", "<u8"),
                            ("method", "<u4"), ("filter", "<u4"), ("crc", "<u4"), ("crc_defined", "u1"), ("props_size", "u1"),
                            ("supported", "u1"), ("pad_", "u1"), ("props", "u1", (8,)), ("first_file", "<u4"),
                            ("num_streams", "<u4")])
SZ_FILE_DTYPE = np.dtype([("size", "<u8"), ("out_off", "<u8"), ("folder", "<u4"), ("crc", "<u4"), ("crc_defined", "u1"),
                          ("is_dir", "u1"), ("has_stream", "u1"), ("attrib_defined", "u1"), ("mtime_defined", "u1"),
                          ("pad_", "u1", (3,)), ("name_off", "<u4"), ("name_len", "<u4"), ("attrib", "<u4"),
                          ("mtime_lo", "<u4"), ("mtime_hi", "<u4"), ("pad2_", "<u4")])
assert SZ_FOLDER_DTYPE.itemsize == 64 and SZ_FILE_DTYPE.itemsize == 56


def sz_extract_all(archive, ctx=None, pinned: bool = False) -> dict:
    """SzArchive_Open + SzArchive_ExtractAll: the reference's SzArEx_Open / SzArEx_Extract for every file.
    pinned: stage archive and output in page-locked memory (what lets the decode pass be a streamed launch)."""
    L = _lib.load()
    arc = _u8(archive)
    if pinned:
        pa = pinned_empty(max(len(arc), 1))
        pa[: len(arc)] = arc
        arc = pa[: len(arc)]
    L.SzArchive_Open.restype = C.c_int
    L.SzArchive_Open.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.SzArchive_Free.argtypes = [C.c_void_p]
    for f in ("SzArchive_NumFolders", "SzArchive_NumFiles"):
        getattr(L, f).restype = C.c_size_t
        getattr(L, f).argtypes = [C.c_void_p]
    for f in ("SzArchive_Folders", "SzArchive_Files"):
        getattr(L, f).restype = C.c_void_p
        getattr(L, f).argtypes = [C.c_void_p]
    L.SzArchive_Names.restype = C.c_void_p
    L.SzArchive_Names.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.SzArchive_UnpackSize.restype = C.c_uint64
    L.SzArchive_UnpackSize.argtypes = [C.c_void_p]
    L.SzArchive_ExtractAll.restype = C.c_int
    L.SzArchive_ExtractAll.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    h = C.c_void_p(None)
    r = L.SzArchive_Open(ctx.h if ctx else None, _p(arc), len(arc), C.byref(h))
    if r == R.SZ_ERROR_FAIL:
        raise _lib.LzbError(_lib.last_error())
    if r != 0:
        return dict(open_res=r, num_files=0, num_folders=0, files=[], folder_res=[], out=np.zeros(0, np.uint8))
    try:
        nfo, nfi = L.SzArchive_NumFolders(h), L.SzArchive_NumFiles(h)
        folders = np.ctypeslib.as_array(C.cast(L.SzArchive_Folders(h), C.POINTER(C.c_uint8)), shape=(max(nfo, 1) * 64,))[: nfo * 64].view(SZ_FOLDER_DTYPE).copy() if nfo else np.zeros(0, SZ_FOLDER_DTYPE)
        files = np.ctypeslib.as_array(C.cast(L.SzArchive_Files(h), C.POINTER(C.c_uint8)), shape=(max(nfi, 1) * 56,))[: nfi * 56].view(SZ_FILE_DTYPE).copy() if nfi else np.zeros(0, SZ_FILE_DTYPE)
        nn = C.c_size_t(0)
        pn = L.SzArchive_Names(h, C.byref(nn))
        names = np.ctypeslib.as_array(C.cast(pn, C.POINTER(C.c_uint16)), shape=(nn.value,)).copy() if nn.value else np.zeros(0, np.uint16)
        total = int(L.SzArchive_UnpackSize(h))
        out = pinned_empty(total + 16) if pinned else np.zeros(total + 16, np.uint8)
        fres, ires = np.zeros(max(nfo, 1), np.int32), np.zeros(max(nfi, 1), np.int32)
        r2 = L.SzArchive_ExtractAll(ctx.h if ctx else None, h, _p(arc), len(arc), _p(out), total, _p(fres), _p(ires))
        if r2 == R.SZ_ERROR_FAIL:
            raise _lib.LzbError(_lib.last_error())
        flist = []
        for k in range(nfi):
            f = files[k]
            nm = names[int(f["name_off"]): int(f["name_off"]) + max(int(f["name_len"]) - 1, 0)].tobytes().decode("utf-16-le") if len(names) else None
            flist.append(dict(res=int(ires[k]), off=int(f["out_off"]), size=int(f["size"]), folder=int(f["folder"]), is_dir=bool(f["is_dir"]),
                              has_stream=bool(f["has_stream"]), crc_defined=bool(f["crc_defined"]), crc=int(f["crc"]), name=nm))
        return dict(open_res=0, extract_res=r2, num_files=nfi, num_folders=nfo, files=flist, folder_res=[int(x) for x in fres[:nfo]],
                    folders=folders, out=out[:total])
    finally:
        L.SzArchive_Free(h)
