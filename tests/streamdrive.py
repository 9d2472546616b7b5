"""Chunked (zlib-like) streaming through the resumable decoder, for three back ends:

  EmuStream : a Python mirror of csrc/lzb_state.cu's marshalling (scalars <-> LzbStreamState, parked bytes
              re-presented in front of the new input, ring dictionary) over the TEST-ONLY emulation of
              the device code -- checks the kernel's resume logic and the re-feed design without a GPU.
  GpuStream : the real C-ABI objects (CLzmaDec / CLzma2Dec + LzmaDec_Allocate / DecodeToBuf ...) of
              liblzb200.so on a GPU.
Both expose lzma_decode_tobuf / lzma2_decode_tobuf with the refbind signature and use the same
feed/drain loop as the reference harness (oracle.refbind.drive_tobuf).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle.refbind import as_u8, drive_tobuf

STATE_DTYPE = np.dtype([
    ("range", "<u4"), ("code", "<u4"), ("state", "<u4"), ("rep", "<u4", (4,)),
    ("total_pos", "<u4"), ("dict_full", "<u4"), ("pending", "<u4"), ("flags", "<u4"), ("dic_pos", "<u4"),
    ("held", "<u4"), ("lc", "<u4"), ("lp", "<u4"), ("pb", "<u4"),
    ("l2_phase", "<u4"), ("l2_control", "<u4"), ("l2_pack_left", "<u4"), ("l2_unpack_left", "<u4"),
    ("l2_flags", "<u4"), ("reserved", "<u4", (11,)),
])
assert STATE_DTYPE.itemsize == 128


class _EmuDecoder:
    """Host half of a resumable decoder, as lzb_state.cu keeps it (fields named like CLzmaDec)."""

    def __init__(self, lc, lp, pb, dict_size, dic_buf_size, lzma2):
        import lzb200
        self.R = lzb200.records
        self.lzma2 = lzma2
        self.lc, self.lp, self.pb, self.dict_size = lc, lp, pb, dict_size
        self.dic = np.zeros(dic_buf_size + 64, np.uint8)
        self.dic_buf_size = dic_buf_size
        mlc, mlp, mpb = (4, 0, 4) if lzma2 else (lc, lp, pb)
        n = 950 + (56 << mpb) + (768 << (mlc + mlp))
        self.blob = np.zeros(128 + 2 * n + 64, np.uint8)       # LzbStreamState + model ("device" memory)
        self.mlc, self.mlp, self.mpb = mlc, mlp, mpb
        self.temp = np.zeros(20, np.uint8)
        self.temp_n = 0
        # LzmaDec_Init / Lzma2Dec_Init
        self.range = self.code = self.state = 0
        self.reps = [0, 0, 0, 0]
        self.processed = self.check_dic = self.remain = 0
        self.need_flush = self.need_init_state = 1
        self.dic_pos = 0
        self.l2 = dict(phase=0, control=0, pack=0, unpack=0, flags=7)

    def to_dic(self, dic_limit, src, finish):
        import emubind
        R = self.R
        st = self.blob[:128].view(STATE_DTYPE)[0]
        st["range"], st["code"], st["state"] = self.range, self.code, self.state
        st["rep"] = self.reps
        st["total_pos"], st["dict_full"], st["pending"] = self.processed, self.check_dic, self.remain
        st["flags"] = self.need_flush | (self.need_init_state << 1)
        st["dic_pos"], st["held"] = self.dic_pos, self.temp_n
        st["lc"], st["lp"], st["pb"] = self.lc, self.lp, self.pb
        st["l2_phase"], st["l2_control"] = self.l2["phase"], self.l2["control"]
        st["l2_pack_left"], st["l2_unpack_left"], st["l2_flags"] = self.l2["pack"], self.l2["unpack"], self.l2["flags"]
        prefix = self.temp_n
        comb = np.concatenate((self.temp[:prefix], src, np.zeros(24, np.uint8)))
        total_in = prefix + len(src)
        u = R.make_units(1)
        u["src_len"], u["dst_cap"], u["dict_size"] = total_in, dic_limit, self.dict_size
        u["codec"] = R.CODEC_LZMA2 if self.lzma2 else R.CODEC_LZMA
        u["lc"], u["lp"], u["pb"], u["finish"] = self.mlc, self.mlp, self.mpb, finish
        u["flags"], u["filter_param"], u["expect"] = 8, self.dic_buf_size, self.blob.ctypes.data
        r = emubind.decode_units(comb, self.dic, u)[0]
        st = self.blob[:128].view(STATE_DTYPE)[0]
        self.range, self.code, self.state = int(st["range"]), int(st["code"]), int(st["state"])
        self.reps = [int(x) for x in st["rep"]]
        self.processed, self.check_dic, self.remain = int(st["total_pos"]), int(st["dict_full"]), int(st["pending"])
        self.need_flush, self.need_init_state = int(st["flags"]) & 1, (int(st["flags"]) >> 1) & 1
        self.dic_pos = int(st["dic_pos"])
        self.lc, self.lp, self.pb = int(st["lc"]), int(st["lp"]), int(st["pb"])
        self.l2 = dict(phase=int(st["l2_phase"]), control=int(st["l2_control"]), pack=int(st["l2_pack_left"]),
                       unpack=int(st["l2_unpack_left"]), flags=int(st["l2_flags"]))
        used, held = int(r["src_used"]), int(st["held"])
        if int(r["status"]) == 3 and held > 0 and used == total_in:
            self.temp[:held] = comb[total_in - held:total_in]
            self.temp_n = held
            src_len = len(src)
        elif used >= prefix:
            self.temp_n = 0
            src_len = used - prefix
        else:
            src_len = 0
        return int(r["res"]), max(int(r["status"]), 0), src_len      # -1 = LZB_ST_INHERIT -> NOT_SPECIFIED

    def to_buf(self, dptr, dl, sptr, sl, finish, st):
        """LzmaDec_DecodeToBuf / Lzma2Dec_DecodeToBuf (reference LzmaDec.c:840-878) over to_dic."""
        out_size, in_size = dl._obj.value, sl._obj.value
        dest = np.ctypeslib.as_array(dptr, shape=(max(out_size, 1),))
        src = np.ctypeslib.as_array(sptr, shape=(max(in_size, 1),))[:in_size]
        made = used = 0
        while True:
            if self.dic_pos == self.dic_buf_size:
                self.dic_pos = 0
            start = self.dic_pos
            if out_size - made > self.dic_buf_size - start:
                stop, mode = self.dic_buf_size, 0
            else:
                stop, mode = start + (out_size - made), finish
            res, status, n_in = self.to_dic(stop, src[used:], mode)
            used += n_in
            got = self.dic_pos - start
            dest[made:made + got] = self.dic[start:start + got]
            made += got
            st._obj.value = status
            if res != 0 or got == 0 or made == out_size:
                dl._obj.value, sl._obj.value = made, used
                return res


class EmuStream:
    kind = "emu-stream"

    def lzma_decode_tobuf(self, packed, props, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        import lzb200
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        try:
            lc, lp, pb, ds = lzb200.records.parse_lzma_props(props)
        except ValueError:
            return dict(res=4, status=0, dest_len=0, src_len=0, out=dst[:dest_cap])
        d = _EmuDecoder(lc, lp, pb, ds, ds, False)
        return drive_tobuf(d.to_buf, src, dst, dest_cap, in_chunk, out_chunk, finish_at_end)

    def lzma2_decode_tobuf(self, packed, prop, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        ds = 0xFFFFFFFF if prop == 40 else (2 | (prop & 1)) << (prop // 2 + 11)
        d = _EmuDecoder(4, 0, 0, ds, ds, True)
        return drive_tobuf(d.to_buf, src, dst, dest_cap, in_chunk, out_chunk, finish_at_end)


# ------------------------------------------------------------------------------------ real C-ABI objects
class CLzmaProps(C.Structure):
    _fields_ = [("lc", C.c_uint), ("lp", C.c_uint), ("pb", C.c_uint), ("dicSize", C.c_uint32)]


class CLzmaDec(C.Structure):
    _fields_ = [("prop", CLzmaProps), ("probs", C.c_void_p), ("dic", C.c_void_p), ("buf", C.c_void_p),
                ("range", C.c_uint32), ("code", C.c_uint32), ("dicPos", C.c_size_t), ("dicBufSize", C.c_size_t),
                ("processedPos", C.c_uint32), ("checkDicSize", C.c_uint32), ("state", C.c_uint),
                ("reps", C.c_uint32 * 4), ("remainLen", C.c_uint), ("needFlush", C.c_int), ("needInitState", C.c_int),
                ("numProbs", C.c_uint32), ("tempBufSize", C.c_uint), ("tempBuf", C.c_uint8 * 20)]


class CLzma2Dec(C.Structure):
    _fields_ = [("decoder", CLzmaDec), ("packSize", C.c_uint32), ("unpackSize", C.c_uint32), ("state", C.c_int),
                ("control", C.c_uint8), ("needInitDic", C.c_int), ("needInitState", C.c_int), ("needInitProp", C.c_int)]


_ALLOC_T = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_FREE_T = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


class ISzAlloc(C.Structure):
    _fields_ = [("Alloc", _ALLOC_T), ("Free", _FREE_T)]


_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]
_alloc = ISzAlloc(_ALLOC_T(lambda p, n: _libc.malloc(n) if n else None), _FREE_T(lambda p, a: _libc.free(a)))


class GpuStream:
    kind = "gpu-stream"

    def __init__(self):
        import lzb200
        L = lzb200.lib.load()
        vp, szp, ip = C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_int)
        L.LzmaDec_Allocate.restype = C.c_int
        L.LzmaDec_Allocate.argtypes = [vp, vp, C.c_uint, vp]
        L.LzmaDec_Free.argtypes = [vp, vp]
        L.LzmaDec_Init.argtypes = [vp]
        L.LzmaDec_DecodeToBuf.restype = C.c_int
        L.LzmaDec_DecodeToBuf.argtypes = [vp, vp, szp, vp, szp, C.c_int, ip]
        L.Lzma2Dec_Allocate.restype = C.c_int
        L.Lzma2Dec_Allocate.argtypes = [vp, C.c_uint8, vp]
        L.Lzma2Dec_Init.argtypes = [vp]
        L.Lzma2Dec_DecodeToBuf.restype = C.c_int
        L.Lzma2Dec_DecodeToBuf.argtypes = [vp, vp, szp, vp, szp, C.c_int, ip]
        self.L = L

    def lzma_decode_tobuf(self, packed, props, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        src, pr = as_u8(packed), as_u8(props)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        d = CLzmaDec()
        r = self.L.LzmaDec_Allocate(C.byref(d), pr.ctypes.data, len(pr), C.byref(_alloc))
        if r != 0:
            return dict(res=r, status=0, dest_len=0, src_len=0, out=dst[:dest_cap])
        self.L.LzmaDec_Init(C.byref(d))
        out = drive_tobuf(lambda *a: self.L.LzmaDec_DecodeToBuf(C.byref(d), *a), src, dst, dest_cap, in_chunk, out_chunk,
                          finish_at_end)
        self.L.LzmaDec_Free(C.byref(d), C.byref(_alloc))
        return out

    def lzma2_decode_tobuf(self, packed, prop, dest_cap, in_chunk, out_chunk, finish_at_end=0):
        src = as_u8(packed)
        dst = np.zeros(max(dest_cap, 1), dtype=np.uint8)
        d = CLzma2Dec()
        r = self.L.Lzma2Dec_Allocate(C.byref(d), prop, C.byref(_alloc))
        if r != 0:
            return dict(res=r, status=0, dest_len=0, src_len=0, out=dst[:dest_cap])
        self.L.Lzma2Dec_Init(C.byref(d))
        out = drive_tobuf(lambda *a: self.L.Lzma2Dec_DecodeToBuf(C.byref(d), *a), src, dst, dest_cap, in_chunk,
                          out_chunk, finish_at_end)
        self.L.LzmaDec_Free(C.byref(d.decoder), C.byref(_alloc))
        return out
