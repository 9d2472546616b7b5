"""ctypes bindings for the two checker libraries -- TEST INFRASTRUCTURE, never product code.

* ``ref()``    -> oracle/_ref/libref.so   : the UNMODIFIED reference compiled from /root/reference
                                             (+ oracle/ref_harness.c).  ``kind = "reference"``.
* ``oracle()`` -> oracle/_build/liboracle.so : this repo's CPU restatement (oracle/*.c). ``kind = "port"``.

Only tests/, tools/ (fixture generation), ``__graft_entry__.smoke()`` and the baseline legs of
``bench.py`` import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref.so")
ORA_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SRC = os.environ.get("LZB_REFERENCE_DIR", "/root/reference")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_szp = C.POINTER(C.c_size_t)
_ip = C.POINTER(C.c_int)


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile liboracle.so, and libref.so when the reference tree is present."""
    kw = dict(stdout=subprocess.DEVNULL) if quiet else {}
    subprocess.check_call(["make", "-s", "-C", HERE], **kw)
    if ref and os.path.isdir(REF_SRC):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={REF_SRC}"], **kw)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _ptr(a: np.ndarray, t=_u8p):
    return a.ctypes.data_as(t)


def as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if len(b) else np.zeros(0, dtype=np.uint8)


_ref = None
_ora = None


class _Ref:
    kind = "reference"

    def __init__(self):
        L = C.CDLL(REF_SO)
        self.L = L
        L.refh_init.restype = None
        L.refh_lzma_encode.restype = C.c_int
        L.refh_lzma_encode.argtypes = [_u8p, C.c_size_t, _u8p, _szp, _u8p, C.c_int, C.c_uint32,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.refh_lzma_decode.restype = C.c_int
        L.refh_lzma_decode.argtypes = [_u8p, _szp, _u8p, _szp, _u8p, C.c_int, _ip]
        L.refh_lzma_decode_tobuf.restype = C.c_int
        L.refh_lzma_decode_tobuf.argtypes = [_u8p, C.c_size_t, _szp, _u8p, C.c_size_t, _szp, _u8p,
                                             C.c_size_t, C.c_size_t, C.c_int, _ip]
        L.refh_lzma2_encode.restype = C.c_int
        L.refh_lzma2_encode.argtypes = [_u8p, C.c_size_t, _u8p, _szp, C.c_int, C.c_uint32,
                                        C.c_int, C.c_int, C.c_int, _u8p]
        L.refh_lzma2_decode.restype = C.c_int
        L.refh_lzma2_decode.argtypes = [_u8p, C.c_size_t, _szp, _u8p, _szp, C.c_uint8, C.c_int, _ip]
        L.refh_lzma2_decode_tobuf.restype = C.c_int
        L.refh_lzma2_decode_tobuf.argtypes = [_u8p, C.c_size_t, _szp, _u8p, C.c_size_t, _szp, C.c_uint8,
                                              C.c_size_t, C.c_size_t, C.c_int, _ip]
        L.refh_decode_batch_mt.restype = C.c_double
        L.refh_decode_batch_mt.argtypes = [C.c_int, C.c_size_t, _u8p, _u64p, _u32p, _u8p, _u64p, _u32p, _u8p,
                                           C.c_int, _i32p, _i32p, _u32p, _u32p, C.c_int, C.c_int]
        L.refh_encode_batch_mt.restype = C.c_double
        L.refh_encode_batch_mt.argtypes = [C.c_int, C.c_size_t, _u8p, _u64p, _u32p, _u8p, _u64p, _u32p, _u32p,
                                           _u8p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _i32p, C.c_int]
        L.refh_xz_assemble.restype = C.c_int
        L.refh_xz_assemble.argtypes = [_u8p, _szp, C.c_size_t, _u8p, _u64p, _u32p, _u32p, _u64p,
                                       C.c_int, C.c_int, C.c_uint32, C.c_uint8]
        L.refh_xz_decode.restype = C.c_int
        L.refh_xz_decode.argtypes = [_u8p, _szp, _u8p, _szp, C.c_int, _ip, _u64p, _ip]
        L.refh_xz_decode_batch_mt.restype = C.c_double
        L.refh_xz_decode_batch_mt.argtypes = [C.c_size_t, _u8p, _u64p, _u32p, _u8p, _u64p, _u32p,
                                              _i32p, _u32p, C.c_int, C.c_int]
        L.refh_x86_convert.restype = C.c_size_t
        L.refh_x86_convert.argtypes = [_u8p, C.c_size_t, C.c_uint32, _u32p, C.c_int]
        L.refh_x86_filter_whole.restype = None
        L.refh_x86_filter_whole.argtypes = [_u8p, C.c_size_t, C.c_uint32, C.c_int]
        L.refh_delta.restype = None
        L.refh_delta.argtypes = [_u8p, C.c_size_t, C.c_uint, C.c_int]
        L.refh_crc32.restype = C.c_uint32
        L.refh_crc32.argtypes = [_u8p, C.c_size_t]
        L.refh_crc64.restype = C.c_uint64
        L.refh_crc64.argtypes = [_u8p, C.c_size_t]
        L.refh_sizeof.restype = C.c_size_t
        L.refh_sizeof.argtypes = [C.c_int]
        L.refh_init()

    # ---- raw LZMA ----
    def lzma_encode(self, data, level=5, dict_size=1 << 16, lc=3, lp=0, pb=2, fb=32, end_mark=0):
        src = as_u8(data)
        cap = len(src) + len(src) // 3 + 1024
        dst = np.empty(cap, dtype=np.uint8)
        dl = C.c_size_t(cap)
        props = np.zeros(5, dtype=np.uint8)
        r = self.L.refh_lzma_encode(_ptr(src), len(src), _ptr(dst), C.byref(dl), _ptr(props), level, dict_size,
                                    lc, lp, pb, fb, end_mark)
        if r != 0:
            raise RuntimeError(f"reference LzmaEncode failed: {r}")
        return dst[: dl.value].copy(), props

    def lzma_decode(self, packed, props, dest_cap, finish_mode=0):
        src = as_u8(packed)
        pr = as_u8(props)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(dest_cap)
        sl = C.c_size_t(len(src))
        st = C.c_int(0)
        r = self.L.refh_lzma_decode(_ptr(dst), C.byref(dl), _ptr(src), C.byref(sl), _ptr(pr), finish_mode, C.byref(st))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])

    def lzma_decode_tobuf(self, packed, props, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        src = as_u8(packed)
        pr = as_u8(props)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(0)
        su = C.c_size_t(0)
        st = C.c_int(0)
        r = self.L.refh_lzma_decode_tobuf(_ptr(dst), dest_cap, C.byref(dl), _ptr(src), len(src), C.byref(su), _ptr(pr),
                                          in_chunk, out_chunk, finish_at_end, C.byref(st))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=su.value, out=dst[:dest_cap])

    # ---- LZMA2 ----
    def lzma2_encode(self, data, level=5, dict_size=1 << 20, lc=-1, lp=-1, pb=-1):
        src = as_u8(data)
        cap = len(src) + len(src) // 8 + 4096
        dst = np.empty(cap, dtype=np.uint8)
        dl = C.c_size_t(cap)
        prop = C.c_uint8(0)
        r = self.L.refh_lzma2_encode(_ptr(src), len(src), _ptr(dst), C.byref(dl), level, dict_size, lc, lp, pb,
                                     C.byref(prop))
        if r != 0:
            raise RuntimeError(f"reference Lzma2Enc_Encode failed: {r}")
        return dst[: dl.value].copy(), prop.value

    def lzma2_decode(self, packed, prop, dest_cap, finish_mode=1):
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(0)
        sl = C.c_size_t(len(src))
        st = C.c_int(0)
        r = self.L.refh_lzma2_decode(_ptr(dst), dest_cap, C.byref(dl), _ptr(src), C.byref(sl), prop, finish_mode,
                                     C.byref(st))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])

    def lzma2_decode_tobuf(self, packed, prop, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(0)
        su = C.c_size_t(0)
        st = C.c_int(0)
        r = self.L.refh_lzma2_decode_tobuf(_ptr(dst), dest_cap, C.byref(dl), _ptr(src), len(src), C.byref(su), prop,
                                           in_chunk, out_chunk, finish_at_end, C.byref(st))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=su.value, out=dst[:dest_cap])

    # ---- batches ----
    def decode_batch_mt(self, kind, src, src_off, src_len, dst_off, dst_cap, props, finish_mode, threads,
                        dst=None, scratch=False):
        n = len(src_len)
        src = as_u8(src)
        src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
        src_len = np.ascontiguousarray(src_len, dtype=np.uint32)
        dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
        dst_cap = np.ascontiguousarray(dst_cap, dtype=np.uint32)
        pr = as_u8(props)
        if dst is None and not scratch:
            total = int((dst_off + dst_cap).max()) if n else 0
            dst = np.zeros(max(total, 1), dtype=np.uint8)
        if dst is None:
            dst = np.zeros(1, dtype=np.uint8)
        res = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        used = np.zeros(n, dtype=np.uint32)
        done = np.zeros(n, dtype=np.uint32)
        secs = self.L.refh_decode_batch_mt(kind, n, _ptr(src), _ptr(src_off, _u64p), _ptr(src_len, _u32p), _ptr(dst),
                                           _ptr(dst_off, _u64p), _ptr(dst_cap, _u32p), _ptr(pr), finish_mode,
                                           _ptr(res, _i32p), _ptr(status, _i32p), _ptr(used, _u32p),
                                           _ptr(done, _u32p), threads, 1 if scratch else 0)
        return dict(secs=secs, res=res, status=status, src_used=used, dst_done=done, out=dst)

    def encode_batch_mt(self, kind, src, src_off, src_len, level, dict_size, lc=3, lp=0, pb=2, fb=32, end_mark=0,
                        threads=1):
        n = len(src_len)
        src = as_u8(src)
        src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
        src_len = np.ascontiguousarray(src_len, dtype=np.uint32)
        cap = (src_len.astype(np.uint64) + src_len.astype(np.uint64) // 3 + 4096).astype(np.uint32)
        dst_off = np.concatenate(([0], np.cumsum(cap.astype(np.uint64))[:-1])).astype(np.uint64)
        dst = np.empty(int(cap.astype(np.uint64).sum()), dtype=np.uint8)
        dlen = np.zeros(n, dtype=np.uint32)
        res = np.zeros(n, dtype=np.int32)
        props = np.zeros(5, dtype=np.uint8)
        secs = self.L.refh_encode_batch_mt(kind, n, _ptr(src), _ptr(src_off, _u64p), _ptr(src_len, _u32p), _ptr(dst),
                                           _ptr(dst_off, _u64p), _ptr(cap, _u32p), _ptr(dlen, _u32p), _ptr(props),
                                           level, dict_size, lc, lp, pb, fb, end_mark, _ptr(res, _i32p), threads)
        if (res != 0).any():
            raise RuntimeError("reference encoder failed on some unit")
        return dict(secs=secs, buf=dst, off=dst_off, len=dlen, props=props if kind == 0 else props[:1])

    # ---- xz ----
    def xz_assemble(self, packed, packed_off, packed_len, unpack_len, check_vals, check_type, filter_id,
                    filter_param, lzma2_prop):
        n = len(packed_len)
        packed = as_u8(packed)
        packed_off = np.ascontiguousarray(packed_off, dtype=np.uint64)
        packed_len = np.ascontiguousarray(packed_len, dtype=np.uint32)
        unpack_len = np.ascontiguousarray(unpack_len, dtype=np.uint32)
        check_vals = np.ascontiguousarray(check_vals, dtype=np.uint64)
        cap = int(packed_len.astype(np.uint64).sum()) + 64 * n + 1024
        out = np.empty(cap, dtype=np.uint8)
        ol = C.c_size_t(cap)
        r = self.L.refh_xz_assemble(_ptr(out), C.byref(ol), n, _ptr(packed), _ptr(packed_off, _u64p),
                                    _ptr(packed_len, _u32p), _ptr(unpack_len, _u32p), _ptr(check_vals, _u64p),
                                    check_type, filter_id, filter_param, lzma2_prop)
        if r != 0:
            raise RuntimeError(f"xz assemble failed: {r}")
        return out[: ol.value].copy()

    def xz_decode(self, data, dest_cap, finish_mode=0):
        src = as_u8(data)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(dest_cap)
        sl = C.c_size_t(len(src))
        st = C.c_int(0)
        nb = C.c_uint64(0)
        fin = C.c_int(0)
        r = self.L.refh_xz_decode(_ptr(dst), C.byref(dl), _ptr(src), C.byref(sl), finish_mode, C.byref(st),
                                  C.byref(nb), C.byref(fin))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, num_blocks=nb.value,
                    finished=bool(fin.value), out=dst[:dest_cap])

    def xz_decode_batch_mt(self, src, src_off, src_len, dst_off, dst_cap, threads, scratch=True):
        n = len(src_len)
        src = as_u8(src)
        src_off = np.ascontiguousarray(src_off, dtype=np.uint64)
        src_len = np.ascontiguousarray(src_len, dtype=np.uint32)
        dst_off = np.ascontiguousarray(dst_off, dtype=np.uint64)
        dst_cap = np.ascontiguousarray(dst_cap, dtype=np.uint32)
        dst = np.zeros(1 if scratch else int((dst_off + dst_cap).max()), dtype=np.uint8)
        res = np.zeros(n, dtype=np.int32)
        done = np.zeros(n, dtype=np.uint32)
        secs = self.L.refh_xz_decode_batch_mt(n, _ptr(src), _ptr(src_off, _u64p), _ptr(src_len, _u32p), _ptr(dst),
                                              _ptr(dst_off, _u64p), _ptr(dst_cap, _u32p), _ptr(res, _i32p),
                                              _ptr(done, _u32p), threads, 1 if scratch else 0)
        return dict(secs=secs, res=res, dst_done=done, out=dst)

    # ---- 7z archive: SzArEx_Open + SzArEx_Extract for every file ----
    def sz_extract_all(self, archive, out_cap=None, max_files=4096):
        arc = as_u8(archive)
        cap = out_cap if out_cap is not None else max(len(arc) * 40, 1 << 20)
        out = np.zeros(cap, dtype=np.uint8)
        nf, nfo = C.c_uint32(0), C.c_uint32(0)
        res = np.zeros(max_files, np.int32); off = np.zeros(max_files, np.uint64); size = np.zeros(max_files, np.uint64)
        folder = np.zeros(max_files, np.uint32); flags = np.zeros(max_files, np.uint8); crc = np.zeros(max_files, np.uint32)
        names = np.zeros(1 << 16, np.uint16)
        nl = C.c_size_t(0)
        f = self.L.refh_7z_extract_all
        f.restype = C.c_int
        f.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_uint32, _u32p, _u32p, _i32p, _u64p, _u64p, _u32p, _u8p,
                      _u32p, C.POINTER(C.c_uint16), C.c_size_t, _szp]
        r = f(_ptr(arc), len(arc), _ptr(out), cap, max_files, C.byref(nf), C.byref(nfo), _ptr(res, _i32p), _ptr(off, _u64p),
              _ptr(size, _u64p), _ptr(folder, _u32p), _ptr(flags), _ptr(crc, _u32p),
              names.ctypes.data_as(C.POINTER(C.c_uint16)), len(names), C.byref(nl))
        n = min(nf.value, max_files)
        text = names[: nl.value].tobytes().decode("utf-16-le").split("\0")[:n] if nl.value else []
        files = [dict(res=int(res[i]), off=int(off[i]), size=int(size[i]), folder=int(folder[i]), is_dir=bool(flags[i] & 1),
                      has_stream=bool(flags[i] & 2), crc_defined=bool(flags[i] & 4), crc=int(crc[i]),
                      name=text[i] if i < len(text) else None) for i in range(n)]
        return dict(open_res=int(r), num_files=nf.value, num_folders=nfo.value, files=files, out=out)

    # ---- filters / checks ----
    def x86_convert(self, data, ip=0, state=0, encoding=0):
        buf = as_u8(data).copy()
        st = C.c_uint32(state)
        n = self.L.refh_x86_convert(_ptr(buf), len(buf), ip, C.byref(st), encoding)
        return buf, int(n), st.value

    def x86_filter_whole(self, data, ip=0, encoding=0):
        buf = as_u8(data).copy()
        self.L.refh_x86_filter_whole(_ptr(buf), len(buf), ip, encoding)
        return buf

    def bra_convert(self, fid, data, ip=0, encoding=0):
        buf = as_u8(data).copy()
        self.L.refh_bra_convert.restype = C.c_size_t
        self.L.refh_bra_convert.argtypes = [C.c_int, _u8p, C.c_size_t, C.c_uint32, C.c_int]
        n = self.L.refh_bra_convert(fid, _ptr(buf), len(buf), ip, encoding)
        return buf, int(n)

    def delta(self, data, dist, encoding=0):
        buf = as_u8(data).copy()
        self.L.refh_delta(_ptr(buf), len(buf), dist, encoding)
        return buf

    def crc32(self, data):
        b = as_u8(data)
        return int(self.L.refh_crc32(_ptr(b), len(b)))

    def crc64(self, data):
        b = as_u8(data)
        return int(self.L.refh_crc64(_ptr(b), len(b)))

    def sizeof(self, which):
        return int(self.L.refh_sizeof(which))


def ref() -> _Ref:
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        _ref = _Ref()
    return _ref


def drive_tobuf(to_buf, src, dst, dest_cap, in_chunk, out_chunk, finish_at_end):
    """The chunked feed/drain loop of refh_lzma_decode_tobuf (oracle/ref_harness.c), over any
    ``to_buf(dest_ptr, &destLen, src_ptr, &srcLen, finishMode, &status)`` callable."""
    ip = op = 0
    st = C.c_int(0)
    n = len(src)
    r = 0
    done = False
    while not done:
        in_left = min(n - ip, in_chunk)
        in_pos = 0
        progressed = False
        while True:
            il = C.c_size_t(in_left)
            ol_req = min(dest_cap - op, out_chunk)
            ol = C.c_size_t(ol_req)
            fm = finish_at_end if ol_req == dest_cap - op else 0
            dptr = dst[op:].ctypes.data_as(_u8p) if op < len(dst) else _ptr(dst)
            sptr = src[ip + in_pos:].ctypes.data_as(_u8p) if ip + in_pos < n else _ptr(src if n else dst)
            r = to_buf(dptr, C.byref(ol), sptr, C.byref(il), fm, C.byref(st))
            in_pos += il.value
            in_left -= il.value
            op += ol.value
            if il.value or ol.value:
                progressed = True
            if r != 0 or st.value == 1 or op == dest_cap:
                done = True
                break
            if in_left == 0 or (il.value == 0 and ol.value == 0):
                break
        ip += in_pos
        if done:
            break
        if ip == n and (not progressed or st.value == 3):
            break
    return dict(res=r, status=st.value, dest_len=op, src_len=ip, out=dst[:dest_cap])


class OraUnit(C.Structure):
    _fields_ = [("src_off", C.c_uint64), ("src_len", C.c_uint64), ("unpack_len", C.c_uint64),
                ("n_chunks", C.c_uint32), ("max_lclp", C.c_uint32)]


class OraXzBlock(C.Structure):
    _fields_ = [("hdr_off", C.c_uint64), ("hdr_size", C.c_uint32), ("data_off", C.c_uint64),
                ("pack_size", C.c_uint64), ("unpack_size", C.c_uint64), ("check_off", C.c_uint64),
                ("n_filters", C.c_int), ("filter_id", C.c_int), ("filter_param", C.c_uint32),
                ("lzma2_prop", C.c_uint8)]


class _OraLzmaState(C.Structure):
    """Opaque storage for OraLzma2 (largest state struct); zero-initialised."""
    _fields_ = [("raw", C.c_uint8 * 512)]


class _Oracle:
    kind = "port"

    def __init__(self):
        L = C.CDLL(ORA_SO)
        self.L = L
        L.ora_lzma_decode.restype = C.c_int
        L.ora_lzma_decode.argtypes = [_u8p, _szp, _u8p, _szp, _u8p, C.c_uint, C.c_int, _ip]
        L.ora_lzma2_decode_unit.restype = C.c_int
        L.ora_lzma2_decode_unit.argtypes = [_u8p, C.c_size_t, _szp, _u8p, _szp, C.c_uint8, C.c_int, _ip]
        L.ora_lzma2_index.restype = C.c_long
        L.ora_lzma2_index.argtypes = [_u8p, C.c_size_t, C.POINTER(OraUnit), C.c_size_t, _szp, _ip]
        L.ora_x86_convert.restype = C.c_size_t
        L.ora_x86_convert.argtypes = [_u8p, C.c_size_t, C.c_uint32, _u32p, C.c_int]
        L.ora_delta_init.restype = None
        L.ora_delta_init.argtypes = [_u8p]
        L.ora_delta_decode.restype = None
        L.ora_delta_decode.argtypes = [_u8p, C.c_uint, _u8p, C.c_size_t]
        L.ora_delta_encode.restype = None
        L.ora_delta_encode.argtypes = [_u8p, C.c_uint, _u8p, C.c_size_t]
        L.ora_crc32.restype = C.c_uint32
        L.ora_crc32.argtypes = [_u8p, C.c_size_t]
        L.ora_crc64.restype = C.c_uint64
        L.ora_crc64.argtypes = [_u8p, C.c_size_t]
        L.ora_xz_index.restype = C.c_int
        L.ora_xz_index.argtypes = [_u8p, C.c_size_t, C.POINTER(OraXzBlock), C.c_size_t, _szp, _ip]
        L.ora_xz_decode.restype = C.c_int
        L.ora_xz_decode.argtypes = [_u8p, _szp, _u8p, _szp, _ip, _u64p, _ip]
        # resumable buffer interface
        L.ora_lzma_alloc.restype = C.c_int
        L.ora_lzma_alloc.argtypes = [C.c_void_p, _u8p, C.c_uint]
        L.ora_lzma_reset.restype = None
        L.ora_lzma_reset.argtypes = [C.c_void_p]
        L.ora_lzma_free.restype = None
        L.ora_lzma_free.argtypes = [C.c_void_p]
        L.ora_lzma_to_buf.restype = C.c_int
        L.ora_lzma_to_buf.argtypes = [C.c_void_p, _u8p, _szp, _u8p, _szp, C.c_int, _ip]
        L.ora_lzma2_alloc.restype = C.c_int
        L.ora_lzma2_alloc.argtypes = [C.c_void_p, C.c_uint8]
        L.ora_lzma2_reset.restype = None
        L.ora_lzma2_reset.argtypes = [C.c_void_p]
        L.ora_lzma2_free.restype = None
        L.ora_lzma2_free.argtypes = [C.c_void_p]
        L.ora_lzma2_to_buf.restype = C.c_int
        L.ora_lzma2_to_buf.argtypes = [C.c_void_p, _u8p, _szp, _u8p, _szp, C.c_int, _ip]

    def lzma_decode(self, packed, props, dest_cap, finish_mode=0):
        src = as_u8(packed)
        pr = as_u8(props)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(dest_cap)
        sl = C.c_size_t(len(src))
        st = C.c_int(0)
        r = self.L.ora_lzma_decode(_ptr(dst), C.byref(dl), _ptr(src), C.byref(sl), _ptr(pr), len(pr), finish_mode,
                                   C.byref(st))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])

    def lzma_decode_tobuf(self, packed, props, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        """Same driver loop as refh_lzma_decode_tobuf, over the oracle's resumable state."""
        src = as_u8(packed)
        pr = as_u8(props)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        stt = _OraLzmaState()
        sp = C.addressof(stt)
        r = self.L.ora_lzma_alloc(sp, _ptr(pr), len(pr))
        if r != 0:
            return dict(res=r, status=0, dest_len=0, src_len=0, out=dst[:dest_cap])
        self.L.ora_lzma_reset(sp)
        out = drive_tobuf(lambda *a: self.L.ora_lzma_to_buf(sp, *a), src, dst, dest_cap, in_chunk, out_chunk, finish_at_end)
        self.L.ora_lzma_free(sp)
        return out

    def lzma2_decode_tobuf(self, packed, prop, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        stt = _OraLzmaState()
        sp = C.addressof(stt)
        r = self.L.ora_lzma2_alloc(sp, prop)
        if r != 0:
            return dict(res=r, status=0, dest_len=0, src_len=0, out=dst[:dest_cap])
        self.L.ora_lzma2_reset(sp)
        out = drive_tobuf(lambda *a: self.L.ora_lzma2_to_buf(sp, *a), src, dst, dest_cap, in_chunk, out_chunk, finish_at_end)
        self.L.ora_lzma2_free(sp)
        return out

    def lzma2_decode(self, packed, prop, dest_cap, finish_mode=1):
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(0)
        sl = C.c_size_t(len(src))
        st = C.c_int(0)
        r = self.L.ora_lzma2_decode_unit(_ptr(dst), dest_cap, C.byref(dl), _ptr(src), C.byref(sl), prop, finish_mode,
                                         C.byref(st))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, out=dst[:dest_cap])

    def lzma2_index(self, packed, max_units=1 << 16):
        src = as_u8(packed)
        units = (OraUnit * max_units)()
        end = C.c_size_t(0)
        eos = C.c_int(0)
        n = self.L.ora_lzma2_index(_ptr(src), len(src), units, max_units, C.byref(end), C.byref(eos))
        out = [dict(src_off=u.src_off, src_len=u.src_len, unpack_len=u.unpack_len, n_chunks=u.n_chunks,
                    max_lclp=u.max_lclp) for u in units[: max(0, min(n, max_units))]]
        return dict(n=n, units=out, end_off=end.value, has_eos=bool(eos.value))

    def x86_convert(self, data, ip=0, state=0, encoding=0):
        buf = as_u8(data).copy()
        st = C.c_uint32(state)
        n = self.L.ora_x86_convert(_ptr(buf), len(buf), ip, C.byref(st), encoding)
        return buf, int(n), st.value

    def x86_filter_whole(self, data, ip=0, encoding=0):
        return self.x86_convert(data, ip, 0, encoding)[0]

    def bra_convert(self, fid, data, ip=0, encoding=0):
        buf = as_u8(data).copy()
        self.L.ora_bra_convert.restype = C.c_size_t
        self.L.ora_bra_convert.argtypes = [C.c_int, _u8p, C.c_size_t, C.c_uint32, C.c_int]
        n = self.L.ora_bra_convert(fid, _ptr(buf), len(buf), ip, encoding)
        return buf, int(n)

    def delta(self, data, dist, encoding=0):
        buf = as_u8(data).copy()
        st = np.zeros(256, dtype=np.uint8)
        (self.L.ora_delta_encode if encoding else self.L.ora_delta_decode)(_ptr(st), dist, _ptr(buf), len(buf))
        return buf

    def crc32(self, data):
        b = as_u8(data)
        return int(self.L.ora_crc32(_ptr(b), len(b)))

    def crc64(self, data):
        b = as_u8(data)
        return int(self.L.ora_crc64(_ptr(b), len(b)))

    def xz_index(self, data, max_blocks=1 << 14):
        src = as_u8(data)
        blocks = (OraXzBlock * max_blocks)()
        nb = C.c_size_t(0)
        ct = C.c_int(0)
        r = self.L.ora_xz_index(_ptr(src), len(src), blocks, max_blocks, C.byref(nb), C.byref(ct))
        out = []
        if r == 0:
            for b in blocks[: nb.value]:
                out.append({f: getattr(b, f) for f, _ in OraXzBlock._fields_})
        return dict(res=r, check_type=ct.value, blocks=out)

    def xz_decode(self, data, dest_cap, finish_mode=0):
        src = as_u8(data)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        dl = C.c_size_t(dest_cap)
        sl = C.c_size_t(len(src))
        st = C.c_int(0)
        nb = C.c_uint64(0)
        fin = C.c_int(0)
        r = self.L.ora_xz_decode(_ptr(dst), C.byref(dl), _ptr(src), C.byref(sl), C.byref(st), C.byref(nb),
                                 C.byref(fin))
        return dict(res=r, status=st.value, dest_len=dl.value, src_len=sl.value, num_blocks=nb.value,
                    finished=bool(fin.value), out=dst[:dest_cap])


def oracle() -> _Oracle:
    global _ora
    if _ora is None:
        if not os.path.exists(ORA_SO):
            build(ref=False)
        _ora = _Oracle()
    return _ora
