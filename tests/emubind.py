"""Loader for the TEST-ONLY host emulation of the device decode logic (tests/emu/emu_core.cpp).

Not a CPU fallback: only tests import this.  It lets `pytest -m "not gpu"` run the kernel's source
(csrc/lzb_core.cuh, lzb_post.cuh compiled with -DLZB_EMU, one emulated lane) against the oracle.
`Emu` exposes the same method names as oracle.refbind so the shared case runner can drive it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "lzma-sdk-zliblike_b200", "csrc")
# LZB_EMU_SANITIZE=1: the same source with AddressSanitizer + UBSan (run python with the sanitizer runtimes preloaded,
# tools/emu_sanitize.sh) -- compute-sanitizer's stand-in for the device code's index arithmetic
SANITIZE = os.environ.get("LZB_EMU_SANITIZE") == "1"
SO = os.path.join(HERE, "emu", "_build", "libemu_san.so" if SANITIZE else "libemu.so")
SRC = os.path.join(HERE, "emu", "emu_core.cpp")

_lib = None
AHEAD = 0          # 1: run the look-ahead form of the decoder (UnitDecoderT<1, false, true>); tests switch it


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    return any(os.path.getmtime(d) > t for d in deps)


def lib():
    global _lib
    if _lib is None:
        if _stale():
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            san = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"] if SANITIZE else []
            subprocess.check_call(["g++", "-O1", "-g", "-shared", "-fPIC", "-std=c++17", "-Wno-unknown-pragmas", *san,
                                   "-I", CSRC, SRC, "-o", SO])
        _lib = C.CDLL(SO)
        _lib.emu_decode_units.restype = C.c_int
        _lib.emu_decode_units.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.emu_decode_units2.restype = C.c_int
        _lib.emu_decode_units2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        _lib.emu_post_units.restype = C.c_int
        _lib.emu_post_units.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    return _lib


def decode_units(packed: np.ndarray, out: np.ndarray, units: np.ndarray) -> np.ndarray:
    import lzb200
    res = np.zeros(len(units), dtype=lzb200.records.RESULT_DTYPE)
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    lib().emu_decode_units2(packed.ctypes.data, out.ctypes.data, units.ctypes.data, res.ctypes.data, len(units), int(AHEAD))
    return res


def post_units(data: np.ndarray, units: np.ndarray) -> np.ndarray:
    import lzb200
    res = np.zeros(len(units), dtype=lzb200.records.RESULT_DTYPE)
    lib().emu_post_use_lanes(0 if AHEAD else 1)        # lean form: per-lane x86 converter; look-ahead form: the tile scan
    lib().emu_post_units(data.ctypes.data, units.ctypes.data, res.ctypes.data, len(units))
    return res


class Emu:
    kind = "emu"

    def _one(self, packed, cap, fill):
        import lzb200
        R = lzb200.records
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        skew = 13                                           # deliberately unaligned unit start
        # under the sanitizers the arenas end where the contract says they may: the packed arena at the end of the last
        # 32-bit word that holds a byte of the unit, the output arena at dst_off + dst_cap (redzones right behind)
        pad = np.zeros((skew + len(packed) + 3) // 4 * 4 if SANITIZE else len(packed) + 64, np.uint8)
        pad[skew:skew + len(packed)] = packed
        u = R.make_units(1)
        u["src_off"], u["src_len"], u["dst_off"], u["dst_cap"] = skew, len(packed), 3, cap
        fill(u, R)
        out = np.zeros(3 + cap if SANITIZE else max(cap, 1) + 64, np.uint8)
        out[:3] = 0xA5
        r = decode_units(pad, out, u)[0]
        # bytes of the arena in front of and behind the unit's slice belong to other units
        assert (out[:3] == 0xA5).all() and not out[3 + cap:].any(), "the unit wrote outside [dst_off, dst_off + dst_cap)"
        # LZB_ST_INHERIT (-1): the unit never assigned *status; a fresh call starts from NOT_SPECIFIED
        return dict(res=int(r["res"]), status=max(int(r["status"]), 0), dest_len=int(r["dst_done"]),
                    src_len=int(r["src_used"]), out=out[3:3 + cap], check=int(r["check"]))

    def lzma_decode(self, packed, props, dest_cap, finish_mode=0):
        import lzb200
        R = lzb200.records
        try:
            lc, lp, pb, ds = R.parse_lzma_props(props)
        except ValueError:
            # host-side check of LzmaDecode (reference LzmaDec.c:981-987): short input first, then props
            short = len(packed) < 5
            return dict(res=R.SZ_ERROR_INPUT_EOF if short else R.SZ_ERROR_UNSUPPORTED, status=0, dest_len=0, src_len=0,
                        out=np.zeros(dest_cap, np.uint8))

        def fill(u, R):
            u["dict_size"], u["codec"], u["lc"], u["lp"], u["pb"], u["finish"] = ds, R.CODEC_LZMA, lc, lp, pb, finish_mode
        return self._one(packed, dest_cap, fill)

    def lzma2_decode(self, packed, prop, dest_cap, finish_mode=1):
        ds = 0xFFFFFFFF if prop == 40 else (2 | (prop & 1)) << (prop // 2 + 11)

        def fill(u, R):
            u["dict_size"], u["codec"], u["finish"] = ds, R.CODEC_LZMA2, finish_mode
        return self._one(packed, dest_cap, fill)

    def _post(self, data, **kw):
        import lzb200
        R = lzb200.records
        buf = np.concatenate((np.full(5, 0xA5, np.uint8), np.ascontiguousarray(data, dtype=np.uint8), np.zeros(0 if SANITIZE else 8, np.uint8)))
        u = R.make_units(1)
        u["dst_off"], u["dst_cap"] = 5, len(data)
        for k, v in kw.items():
            u[k] = v
        r = post_units(buf, u)[0]
        assert (buf[:5] == 0xA5).all() and not buf[5 + len(data):].any(), "the pass wrote outside the unit's bytes"
        return buf[5:5 + len(data)], r

    def x86_convert(self, data, ip=0, state=0, encoding=0):
        if len(data) < 5:
            return np.ascontiguousarray(data, dtype=np.uint8).copy(), 0, state
        out, r = self._post(data, filter=4, filter_param=ip, flags=4 | (2 if encoding else 0), expect=(state & 7) << 32)
        return out, int(r["src_used"]), int(r["status"])

    def delta(self, data, dist, encoding=0):
        return self._post(data, filter=3, filter_param=dist, flags=2 if encoding else 0)[0]

    def bra_convert(self, fid, data, ip=0, encoding=0):
        if len(data) < (16 if fid == 6 else 4):
            return np.ascontiguousarray(data, dtype=np.uint8).copy(), 0
        out, r = self._post(data, filter=fid, filter_param=ip, flags=4 | (2 if encoding else 0))
        return out, int(r["src_used"])

    def crc32(self, data):
        return int(self._post(data, check=1)[1]["check"])

    def crc64(self, data):
        return int(self._post(data, check=4)[1]["check"])
