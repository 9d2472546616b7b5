"""Minimal 7z archive writer -- TEST INFRASTRUCTURE.

The reference can read 7z archives (7zIn.c / 7zDec.c) but has no writer, so fixtures for the folder path are
assembled here from payloads compressed by the reference's own encoders.  Written per the 7z format as the
reference's reader consumes it (7zIn.c:346-1110): signature header, packed streams, header (optionally LZMA-
packed: kEncodedHeader) with PackInfo / UnpackInfo / SubStreamsInfo / FilesInfo.

Folder shapes (what SzFolder_Decode accepts, 7zDec.c:269-331): one coder (Copy / LZMA / LZMA2), or a main
coder followed by BCJ x86 or ARM.  `delta` as second coder can be written too: the reference answers
SZ_ERROR_UNSUPPORTED for it (7zDec.c:291-298), which is a parity case.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

K_END, K_HEADER, K_MAIN_STREAMS, K_FILES, K_PACK, K_UNPACK, K_SUBSTREAMS, K_SIZE, K_CRC, K_FOLDER = 0, 1, 4, 5, 6, 7, 8, 9, 10, 11
K_CODERS_UNPACK, K_NUM_UNPACK, K_EMPTY_STREAM, K_EMPTY_FILE, K_NAME, K_MTIME, K_ATTR, K_ENCODED = 12, 13, 14, 15, 17, 20, 21, 23
SIG = b"7z\xBC\xAF\x27\x1C"
M_COPY, M_LZMA, M_LZMA2 = b"\x00", b"\x03\x01\x01", b"\x21"
M_BCJ, M_ARM, M_DELTA = b"\x03\x03\x01\x03", b"\x03\x03\x05\x01", b"\x03"


def num(v: int) -> bytes:
    """7z variable-length number (7zIn.c:346-366): leading 1-bits of the first byte count extra bytes."""
    for extra in range(9):
        if v < (1 << (7 * (extra + 1))) or extra == 8:
            if extra == 8:
                return b"\xFF" + v.to_bytes(8, "little")
            first = ((0xFF << (8 - extra)) & 0xFF) | (v >> (8 * extra))
            return bytes([first]) + (v & ((1 << (8 * extra)) - 1)).to_bytes(extra, "little")
    raise AssertionError


def bits(flags) -> bytes:
    out, cur, n = bytearray(), 0, 0
    for f in flags:
        cur = (cur << 1) | (1 if f else 0)
        n += 1
        if n == 8:
            out.append(cur); cur = n = 0
    if n:
        out.append(cur << (8 - n))
    return bytes(out)


@dataclass
class Folder:
    method: str                      # "copy" | "lzma" | "lzma2"
    files: list                      # [(name, bytes)] in stream order
    filt: str | None = None          # None | "bcj" | "arm" | "delta"
    level: int = 1
    dict_size: int = 1 << 16
    folder_crc: bool = True
    file_crc: bool = True
    packed: bytes = b""
    props: bytes = b""
    plain: bytes = b""


def _filter(ref, data: bytes, filt):
    import numpy as np
    a = np.frombuffer(data, dtype=np.uint8)
    if filt == "bcj":
        return ref.x86_filter_whole(a, 0, 1).tobytes()
    if filt == "arm":
        return ref.bra_convert(7, a, 0, 1)[0].tobytes()
    if filt == "delta":
        return ref.delta(a, 1, 1).tobytes()
    return data


def encode_folder(ref, f: Folder):
    import numpy as np
    f.plain = b"".join(d for _, d in f.files)
    pre = _filter(ref, f.plain, f.filt)
    if f.method == "copy":
        f.packed, f.props = pre, b""
    elif f.method == "lzma":
        pk, props = ref.lzma_encode(np.frombuffer(pre, dtype=np.uint8), level=f.level, dict_size=f.dict_size)
        f.packed, f.props = pk.tobytes(), props.tobytes()
    else:
        pk, prop = ref.lzma2_encode(np.frombuffer(pre, dtype=np.uint8), level=f.level, dict_size=max(f.dict_size, 1 << 12))
        f.packed, f.props = pk.tobytes(), bytes([prop])
    return f


def _coder(method: bytes, props: bytes = b"") -> bytes:
    main = len(method) | (0x20 if props else 0)
    return bytes([main]) + method + (num(len(props)) + props if props else b"")


def _folder_record(f: Folder) -> bytes:
    main = {"copy": M_COPY, "lzma": M_LZMA, "lzma2": M_LZMA2}[f.method]
    if f.filt is None:
        return num(1) + _coder(main, f.props)
    second = {"bcj": M_BCJ, "arm": M_ARM, "delta": M_DELTA}[f.filt]
    sprops = b"\x00" if f.filt == "delta" else b""
    # coder 0 = main (in 0, out 0), coder 1 = filter (in 1, out 1); bind pair: in 1 <- out 0
    return num(2) + _coder(main, f.props) + _coder(second, sprops) + num(1) + num(0)


def _streams_info(folders, data_offset: int, sub=True) -> bytes:
    live = [f for f in folders]
    out = bytearray([K_PACK]) + num(data_offset) + num(len(live)) + bytes([K_SIZE])
    for f in live:
        out += num(len(f.packed))
    out += bytes([K_END])
    out += bytes([K_UNPACK, K_FOLDER]) + num(len(live)) + b"\x00"
    for f in live:
        out += _folder_record(f)
    out += bytes([K_CODERS_UNPACK])
    for f in live:
        out += num(len(f.plain))
        if f.filt is not None:
            out += num(len(f.plain))
    if any(f.folder_crc for f in live):
        alld = all(f.folder_crc for f in live)
        out += bytes([K_CRC]) + (b"\x01" if alld else b"\x00" + bits(f.folder_crc for f in live))
        for f in live:
            if f.folder_crc:
                out += struct.pack("<I", zlib.crc32(f.plain))
    out += bytes([K_END])
    if sub:
        out += bytes([K_SUBSTREAMS])
        if any(len(f.files) != 1 for f in live):
            out += bytes([K_NUM_UNPACK])
            for f in live:
                out += num(len(f.files))
            if any(len(f.files) > 1 for f in live):
                out += bytes([K_SIZE])
                for f in live:
                    for _, d in f.files[:-1]:
                        out += num(len(d))
        # digests for every substream of folders that have != 1 stream or no folder CRC
        need = [(f, d) for f in live for _, d in f.files if len(f.files) != 1 or not f.folder_crc]
        if need and any(f.file_crc for f, _ in need):
            alld = all(f.file_crc for f, _ in need)
            out += bytes([K_CRC]) + (b"\x01" if alld else b"\x00" + bits(f.file_crc for f, _ in need))
            for f, d in need:
                if f.file_crc:
                    out += struct.pack("<I", zlib.crc32(d))
        out += bytes([K_END])
    out += bytes([K_END])
    return bytes(out)


def write_archive(ref, folders, empty_files=(), dirs=(), encode_header=False, mtime=True, junk_after=0) -> bytes:
    """folders: [Folder] (encoded here).  Files are listed folder by folder, then empty files, then dirs."""
    for f in folders:
        if not f.packed and not f.plain:
            encode_folder(ref, f)
    packed_area = b"".join(f.packed for f in folders)
    names = [n for f in folders for n, _ in f.files] + list(empty_files) + list(dirs)
    nfiles = len(names)
    nstream = sum(len(f.files) for f in folders)
    hdr = bytearray([K_HEADER])
    if folders:
        hdr += bytes([K_MAIN_STREAMS]) + _streams_info(folders, 0)
    hdr += bytes([K_FILES]) + num(nfiles)
    if nfiles != nstream:
        v = bits([i >= nstream for i in range(nfiles)])
        hdr += bytes([K_EMPTY_STREAM]) + num(len(v)) + v
        if empty_files:
            ne = nfiles - nstream
            v = bits([i < len(empty_files) for i in range(ne)])
            hdr += bytes([K_EMPTY_FILE]) + num(len(v)) + v
    nb = b"".join(n.encode("utf-16-le") + b"\0\0" for n in names)
    hdr += bytes([K_NAME]) + num(len(nb) + 1) + b"\x00" + nb
    if mtime:
        body = b"\x01\x00" + b"".join(struct.pack("<II", 0x1234 + i, 0x01D0) for i in range(nfiles))
        hdr += bytes([K_MTIME]) + num(len(body)) + body
        body = b"\x01\x00" + b"".join(struct.pack("<I", 0x10 if i >= nfiles - len(dirs) else 0x20) for i in range(nfiles))
        hdr += bytes([K_ATTR]) + num(len(body)) + body
    hdr += bytes([K_END, K_END])
    hdr = bytes(hdr)
    tail = hdr
    if encode_header:
        import numpy as np
        pk, props = ref.lzma_encode(np.frombuffer(hdr, dtype=np.uint8), level=5, dict_size=1 << 16)
        hf = Folder("lzma", [("", hdr)], folder_crc=True)
        hf.packed, hf.props, hf.plain = pk.tobytes(), props.tobytes(), hdr
        tail = bytes([K_ENCODED]) + _streams_info([hf], len(packed_area), sub=False)
        packed_area += hf.packed
    start = struct.pack("<QQI", len(packed_area), len(tail), zlib.crc32(tail))
    sig = SIG + bytes([0, 3]) + struct.pack("<I", zlib.crc32(start)) + start
    return sig + packed_area + tail + b"\x5A" * junk_after


def expected_listing(folders, empty_files=(), dirs=()):
    """[(name, bytes or None for a directory)] in archive order."""
    out = [(n, d) for f in folders for n, d in f.files]
    out += [(n, b"") for n in empty_files] + [(n, None) for n in dirs]
    return out
