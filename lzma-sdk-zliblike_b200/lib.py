"""ctypes loader for lib/liblzb200.so.  Fails loudly: no library or no CUDA device -> exception."""
from __future__ import annotations

import ctypes as C
import os
import re

from . import buildlib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lzb200.h")

_u8p = C.POINTER(C.c_uint8)
_szp = C.POINTER(C.c_size_t)
_ip = C.POINTER(C.c_int)

_lib = None


class LzbError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function / data symbol include/lzb200.h declares (parsed from the header text)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = "\n".join(l for l in txt.splitlines() if not l.strip().startswith("#"))
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt))
    names |= set(re.findall(r"extern\s+\w+\s+(\w+)\[\]", txt))
    names -= {"Alloc", "Free", "sizeof", "SRes", "void"}      # function-pointer fields, keywords
    return sorted(n for n in names if not n.startswith("_"))


def load(build: bool = True) -> C.CDLL:
    """dlopen the library (building it first when stale).  Loading does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    so = buildlib.build() if build else buildlib.SO
    if not os.path.exists(so):
        raise LzbError(f"{so} is missing: build it with lzma-sdk-zliblike_b200/buildlib.py (there is no CPU fallback)")
    L = C.CDLL(so)
    L.Lzb_LastError.restype = C.c_char_p
    L.Lzb_Create.restype = C.c_void_p
    L.Lzb_Create.argtypes = [C.c_int]
    L.Lzb_Destroy.argtypes = [C.c_void_p]
    L.Lzb_DeviceSMs.argtypes = [C.c_void_p]
    L.Lzb_HostAlloc.restype = C.c_void_p
    L.Lzb_HostAlloc.argtypes = [C.c_size_t]
    L.Lzb_HostFree.argtypes = [C.c_void_p]
    L.Lzb_DecodeUnitsHost.restype = C.c_int
    L.Lzb_DecodeUnitsHost.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_void_p, C.c_size_t]
    L.Lzb_DecodeUnitsDevice.restype = C.c_int
    L.Lzb_DecodeUnitsDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_uint, C.c_uint, C.c_void_p]
    L.LzmaUncompress.restype = C.c_int
    L.LzmaUncompress.argtypes = [C.c_void_p, _szp, C.c_void_p, _szp, C.c_void_p, C.c_size_t]
    L.LzmaDecode.restype = C.c_int
    L.LzmaDecode.argtypes = [C.c_void_p, _szp, C.c_void_p, _szp, C.c_void_p, C.c_uint, C.c_int, _ip, C.c_void_p]
    L.LzmaUncompressBatch.restype = C.c_int
    L.LzmaUncompressBatch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_void_p]
    L.Lzma2_IndexUnits.restype = C.c_int
    L.Lzma2_IndexUnits.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p, C.c_size_t, _szp, _szp, _ip]
    L.Lzma2DecodeParallel.restype = C.c_int
    L.Lzma2DecodeParallel.argtypes = [C.c_void_p, C.c_void_p, _szp, C.c_void_p, _szp, C.c_uint8, C.c_int, _ip]
    L.Lzma2Decode.restype = C.c_int
    L.Lzma2Decode.argtypes = [C.c_void_p, _szp, C.c_void_p, _szp, C.c_uint8, C.c_int, _ip, C.c_void_p]
    L.XzDecodeParallel.restype = C.c_int
    L.XzDecodeParallel.argtypes = [C.c_void_p, C.c_void_p, _szp, C.c_void_p, _szp, _ip, C.POINTER(C.c_uint64), _ip]
    L.Xz_ReadIndexBackward.restype = C.c_int
    L.Xz_ReadIndexBackward.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, _szp, _ip]
    L.CrcCalc.restype = C.c_uint32
    L.CrcCalc.argtypes = [C.c_void_p, C.c_size_t]
    L.CrcUpdate.restype = C.c_uint32
    L.CrcUpdate.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
    L.Crc64Calc.restype = C.c_uint64
    L.Crc64Calc.argtypes = [C.c_void_p, C.c_size_t]
    L.Crc64Update.restype = C.c_uint64
    L.Crc64Update.argtypes = [C.c_uint64, C.c_void_p, C.c_size_t]
    L.x86_Convert.restype = C.c_size_t
    L.x86_Convert.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32), C.c_int]
    L.Delta_Init.argtypes = [C.c_void_p]
    L.Delta_Decode.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_size_t]
    L.LzmaProps_Decode.restype = C.c_int
    L.LzmaProps_Decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    _lib = L
    return L


def last_error() -> str:
    return (load().Lzb_LastError() or b"").decode()


class Context:
    """One GPU context (Lzb_Create).  Raises when there is no CUDA device."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = self.L.Lzb_Create(device)
        if not self.h:
            raise LzbError(f"Lzb_Create({device}) failed: {last_error()}")

    def close(self):
        if self.h:
            self.L.Lzb_Destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sms(self) -> int:
        return int(self.L.Lzb_DeviceSMs(self.h))
