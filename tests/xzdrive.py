"""Chunked driver for XzUnpacker_Code, over ANY shared library that exports the SDK's xz symbols with the
reference ABI: oracle/_ref/libref.so (the unmodified reference) or liblzb200.so (the product).  Every call's
(SRes, status, srcLen, destLen) goes into a trace, so two implementations can be compared call by call.
"""
from __future__ import annotations

import ctypes as C
import zlib

import numpy as np

XZ_UNPACKER_SIZE = 1792          # sizeof(CXzUnpacker), x86-64
OFF_NUM_BLOCKS = 40              # offsetof(CXzUnpacker, numBlocks)

_ALLOC_T = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_FREE_T = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]


class ISzAlloc(C.Structure):
    _fields_ = [("Alloc", _ALLOC_T), ("Free", _FREE_T)]


def _mk_alloc():
    a = ISzAlloc(_ALLOC_T(lambda p, n: _libc.malloc(n) if n else None), _FREE_T(lambda p, a: _libc.free(a)))
    return a


def bind(L):
    L.XzUnpacker_Create.restype = C.c_int
    L.XzUnpacker_Create.argtypes = [C.c_void_p, C.c_void_p]
    L.XzUnpacker_Free.restype = None
    L.XzUnpacker_Free.argtypes = [C.c_void_p]
    L.XzUnpacker_Code.restype = C.c_int
    L.XzUnpacker_Code.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.POINTER(C.c_size_t),
                                  C.c_int, C.POINTER(C.c_int)]
    L.XzUnpacker_IsStreamWasFinished.restype = C.c_int
    L.XzUnpacker_IsStreamWasFinished.argtypes = [C.c_void_p]
    L.CrcGenerateTable.restype = None
    L.Crc64GenerateTable.restype = None
    L.CrcGenerateTable()
    L.Crc64GenerateTable()
    return L


def drive_xz(L, data, dest_cap, in_chunk, out_chunk, finish_mode=0, max_calls=2_000_000):
    """Feed `data` in_chunk bytes at a time, drain at most out_chunk bytes per call, until the decoder stops
    making progress with all input presented, fails, or the output capacity is used up and a call with no
    room makes no progress.  Returns the per-call trace and the totals."""
    src = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8)) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    n = len(src)
    srcbuf = np.concatenate((src, np.zeros(8, np.uint8)))
    dst = np.zeros(dest_cap + 8, np.uint8)
    obj = C.create_string_buffer(XZ_UNPACKER_SIZE)
    alloc = _mk_alloc()
    L.XzUnpacker_Create(obj, C.byref(alloc))
    trace = []
    ip = op = 0
    res, st = 0, C.c_int(0)
    calls = 0
    stop = False
    while not stop:
        fed = min(in_chunk, n - ip)
        used = 0
        while True:
            il = C.c_size_t(fed - used)
            ol = C.c_size_t(min(out_chunk, dest_cap - op))
            res = L.XzUnpacker_Code(obj, dst.ctypes.data + op, C.byref(ol), srcbuf.ctypes.data + ip + used, C.byref(il),
                                    finish_mode, C.byref(st))
            calls += 1
            trace.append((res, st.value, il.value, ol.value))
            used += il.value
            op += ol.value
            if res != 0 or calls >= max_calls:
                stop = True
                break
            if il.value == 0 and ol.value == 0:
                break
        ip += used
        if stop:
            break
        if used < fed or ip >= n:
            break        # input refused (output full / decoder stuck), or everything presented and drained
    finished = bool(L.XzUnpacker_IsStreamWasFinished(obj))
    num_blocks = int(np.frombuffer(obj.raw[OFF_NUM_BLOCKS:OFF_NUM_BLOCKS + 8], dtype="<u8")[0])
    L.XzUnpacker_Free(obj)
    t = np.asarray(trace, dtype=np.int64).reshape(-1, 4)
    return dict(res=int(res), status=int(st.value), src_len=int(ip), dest_len=int(op), calls=int(calls),
                trace=t, trace_crc=int(zlib.crc32(t.tobytes())), finished=finished, num_blocks=num_blocks,
                out=dst[:op].copy(), out_crc=int(zlib.crc32(dst[:op].tobytes())))


def summary(d):
    """What the golden file keeps of a drive."""
    return {k: d[k] for k in ("res", "status", "src_len", "dest_len", "calls", "trace_crc", "finished", "num_blocks", "out_crc")}
