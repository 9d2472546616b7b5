"""numpy views of the C records in csrc/lzb_types.h (LzbUnit / LzbResult) and the SDK's codes."""
from __future__ import annotations

import numpy as np

UNIT_DTYPE = np.dtype([
    ("src_off", "<u8"), ("dst_off", "<u8"), ("src_len", "<u4"), ("dst_cap", "<u4"), ("dict_size", "<u4"),
    ("codec", "u1"), ("lc", "u1"), ("lp", "u1"), ("pb", "u1"),
    ("finish", "u1"), ("filter", "u1"), ("check", "u1"), ("flags", "u1"),
    ("filter_param", "<u4"), ("expect", "<u8"),
])
RESULT_DTYPE = np.dtype([
    ("res", "<i4"), ("status", "<i4"), ("src_used", "<u4"), ("dst_done", "<u4"), ("check", "<u8"),
])
assert UNIT_DTYPE.itemsize == 48 and RESULT_DTYPE.itemsize == 24

CODEC_LZMA, CODEC_LZMA2 = 0, 1
FILTER_NONE, FILTER_DELTA, FILTER_X86 = 0, 3, 4
CHECK_NONE, CHECK_CRC32, CHECK_CRC64 = 0, 1, 4
UNIT_VERIFY = 1

# SRes (reference Types.h:25-41)
SZ_OK, SZ_ERROR_DATA, SZ_ERROR_MEM, SZ_ERROR_CRC, SZ_ERROR_UNSUPPORTED = 0, 1, 2, 3, 4
SZ_ERROR_PARAM, SZ_ERROR_INPUT_EOF, SZ_ERROR_OUTPUT_EOF = 5, 6, 7
SZ_ERROR_FAIL, SZ_ERROR_ARCHIVE, SZ_ERROR_NO_ARCHIVE = 11, 16, 17
# ELzmaStatus (reference LzmaDec.h:100-107)
LZMA_STATUS_NOT_SPECIFIED, LZMA_STATUS_FINISHED_WITH_MARK, LZMA_STATUS_NOT_FINISHED = 0, 1, 2
LZMA_STATUS_NEEDS_MORE_INPUT, LZMA_STATUS_MAYBE_FINISHED_WITHOUT_MARK = 3, 4
# ELzmaFinishMode (reference LzmaDec.h:79-83)
LZMA_FINISH_ANY, LZMA_FINISH_END = 0, 1


def parse_lzma_props(props) -> tuple[int, int, int, int]:
    """(lc, lp, pb, dict_size) from the 5 property bytes; raises ValueError like SZ_ERROR_UNSUPPORTED.

    Mirrors LzmaProps_Decode (reference LzmaDec.c:898-922): dictionary clamped up to 4096.
    """
    p = bytes(props)
    if len(p) < 5 or p[0] >= 225:
        raise ValueError("unsupported LZMA properties")
    d = p[0]
    lc, d = d % 9, d // 9
    lp, pb = d % 5, d // 5
    ds = int.from_bytes(p[1:5], "little")
    return lc, lp, pb, max(ds, 1 << 12)


def make_units(n: int) -> np.ndarray:
    return np.zeros(n, dtype=UNIT_DTYPE)
