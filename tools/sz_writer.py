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
M_BCJ, M_ARM, M_DELTA, M_BCJ2 = b"\x03\x03\x01\x03", b"\x03\x03\x05\x01", b"\x03", b"\x03\x03\x01\x1B"


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
    # filt == "bcj2": the four pack streams [main, range coder, call, jump], the props of the three LZMA coders
    # [jump, call, main] (coder order) and the sizes of their outputs
    pack_list: list = field(default_factory=list)
    sub_props: list = field(default_factory=list)
    sub_sizes: list = field(default_factory=list)


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


class _RangeEnc:
    """LZMA-style binary range encoder (the inverse of the decoder in Bcj2.c:9-21)."""
    def __init__(self):
        self.low, self.range, self.cache, self.cache_size, self.out = 0, 0xFFFFFFFF, 0, 1, bytearray()

    def _shift(self):
        if self.low < 0xFF000000 or self.low >= (1 << 32):
            carry = self.low >> 32
            temp = self.cache
            while True:
                self.out.append((temp + carry) & 0xFF)
                temp = 0xFF
                self.cache_size -= 1
                if self.cache_size == 0:
                    break
            self.cache = (self.low >> 24) & 0xFF
        self.cache_size += 1
        self.low = (self.low & 0x00FFFFFF) << 8

    def bit(self, probs, i, b):
        p = probs[i]
        bound = (self.range >> 11) * p
        if b == 0:
            self.range = bound
            probs[i] = p + ((2048 - p) >> 5)
        else:
            self.low += bound
            self.range -= bound
            probs[i] = p - (p >> 5)
        while self.range < (1 << 24):
            self.range = (self.range << 8) & 0xFFFFFFFF
            self._shift()

    def finish(self):
        for _ in range(5):
            self._shift()
        return bytes(self.out)


def bcj2_encode(data: bytes, every: int = 3):
    """Split x86 code into (main, call, jump, rc) as the decoder of Bcj2.c:28-127 expects them.  Which branch
    targets are moved is the encoder's choice; here: all whose four operand bytes exist, except every `every`-th
    candidate (so that both decisions and the in-operand rescans are exercised)."""
    n = len(data)
    main, call, jump = bytearray(), bytearray(), bytearray()
    rc, probs = _RangeEnc(), [1024] * 258
    prev, i, k = 0, 0, 0
    while i < n:
        b = data[i]
        main.append(b)
        is_j = (b & 0xFE) == 0xE8 or (prev == 0x0F and (b & 0xF0) == 0x80)
        i += 1
        if not is_j:
            prev = b
            continue
        if i == n:                      # the opcode is the last byte: the decoder stops before asking for a bit
            break
        ctx = prev if b == 0xE8 else (256 if b == 0xE9 else 257)
        k += 1
        if i + 4 <= n and k % every != 0:
            rc.bit(probs, ctx, 1)
            rel = int.from_bytes(data[i:i + 4], "little")
            dest = (rel + i + 4) & 0xFFFFFFFF
            (call if b == 0xE8 else jump).extend(dest.to_bytes(4, "big"))
            prev = data[i + 3]
            i += 4
        else:
            rc.bit(probs, ctx, 0)
            prev = b
    return bytes(main), bytes(call), bytes(jump), rc.finish()


def encode_folder(ref, f: Folder):
    import numpy as np
    f.plain = b"".join(d for _, d in f.files)
    if f.filt == "bcj2":
        main, call, jump, rcs = bcj2_encode(f.plain)
        packs, props = [], []
        for part in (jump, call, main):                  # coder order 0, 1, 2
            pk, pr = ref.lzma_encode(np.frombuffer(part, dtype=np.uint8) if part else np.zeros(0, np.uint8), level=f.level, dict_size=f.dict_size)
            packs.append(pk.tobytes()); props.append(pr.tobytes())
        f.pack_list = [packs[2], rcs, packs[1], packs[0]]  # pack stream order: main, range coder, call, jump
        f.sub_props, f.sub_sizes = props, [len(jump), len(call), len(main)]
        f.packed = b"".join(f.pack_list)
        return f
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
    if f.filt == "bcj2":
        # three LZMA coders (in 0,1,2 / out 0,1,2), then BCJ2 with four inputs (in 3..6) and one output (out 3);
        # its inputs 3,4,5 are fed by outputs 2,1,0; pack streams feed inputs 2, 6, 1, 0 (7zDec.c:308-329)
        rec = num(4)
        for pr in f.sub_props:
            rec += _coder(M_LZMA, pr)
        rec += bytes([len(M_BCJ2) | 0x10]) + M_BCJ2 + num(4) + num(1)
        for a, b in ((5, 0), (4, 1), (3, 2)):
            rec += num(a) + num(b)
        for ps in (2, 6, 1, 0):
            rec += num(ps)
        return rec
    main = {"copy": M_COPY, "lzma": M_LZMA, "lzma2": M_LZMA2}[f.method]
    if f.filt is None:
        return num(1) + _coder(main, f.props)
    second = {"bcj": M_BCJ, "arm": M_ARM, "delta": M_DELTA}[f.filt]
    sprops = b"\x00" if f.filt == "delta" else b""
    # coder 0 = main (in 0, out 0), coder 1 = filter (in 1, out 1); bind pair: in 1 <- out 0
    return num(2) + _coder(main, f.props) + _coder(second, sprops) + num(1) + num(0)


def _streams_info(folders, data_offset: int, sub=True) -> bytes:
    live = [f for f in folders]
    streams = [p for f in live for p in (f.pack_list if f.pack_list else [f.packed])]
    out = bytearray([K_PACK]) + num(data_offset) + num(len(streams)) + bytes([K_SIZE])
    for p in streams:
        out += num(len(p))
    out += bytes([K_END])
    out += bytes([K_UNPACK, K_FOLDER]) + num(len(live)) + b"\x00"
    for f in live:
        out += _folder_record(f)
    out += bytes([K_CODERS_UNPACK])
    for f in live:
        if f.filt == "bcj2":
            for z in f.sub_sizes:
                out += num(z)
            out += num(len(f.plain))
            continue
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
