"""Golden vectors for the stateless branch converters (ARM, ARMT, PPC, SPARC, IA64) and Delta_Encode:
outputs of the unmodified reference (oracle/_ref/libref.so) on seeded synthetic "code".

    python tools/make_golden_bra.py   ->  tests/golden/bra.json + bra.npz
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")
IDS = {"ppc": 5, "ia64": 6, "arm": 7, "armt": 8, "sparc": 9}


def synth(arch: str, n: int, seed: int) -> np.ndarray:
    """Random bytes with branch-and-link opcodes of `arch` planted every few slots."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, n, dtype=np.uint8)
    if arch == "arm":
        d[3::28] = 0xEB
    elif arch == "ppc":
        d[0::20] = 0x48 | (d[0::20] & 3)
        d[3::20] = (d[3::20] & 0xFC) | 1
    elif arch == "sparc":
        d[0::24] = 0x40
        d[1::24] &= 0x3F
        d[12::48] = 0x7F
        d[13::48] |= 0xC0
    elif arch == "armt":
        d[1::18] = 0xF0 | (d[1::18] & 7)
        d[3::18] = 0xF8 | (d[3::18] & 7)
        d[101::36] = 0xF4                       # a BL prefix whose partner is missing / overlapping patterns
    elif arch == "ia64":
        d[0::16] = (d[0::16] & 0xE0) | rng.choice(np.array([16, 17, 18, 19, 22, 23, 24, 25, 28, 29, 3, 11], dtype=np.uint8), size=len(d[0::16]))
        for b in range(0, n - 16, 48):          # make slot 2 of every third bundle a br.call (opcode 5, btype 0)
            v = int.from_bytes(d[b:b + 16].tobytes(), "little")
            s = 5 + 82
            v &= ~(((0xF << 37) | (0x7 << 9)) << s)
            v |= (0x5 << 37) << s
            d[b:b + 16] = np.frombuffer(v.to_bytes(16, "little"), dtype=np.uint8)
    return d


def main():
    from oracle import refbind
    ref = refbind.ref()
    blobs, cases = {}, []
    for arch, fid in IDS.items():
        blob = synth(arch, 6000 + 37 * fid, seed=100 + fid)
        blobs[arch] = blob
        for n in (0, 3, 4, 5, 15, 16, 17, 31, 33, 1000, 1001, 1002, 1003, len(blob)):
            for ip, enc in ((0, 0), (0, 1), (0x10000, 0), (0xFFFFF000, 1)):
                out, done = ref.bra_convert(fid, blob[:n], ip, enc)
                cases.append(dict(kind="bra", blob=arch, id=fid, n=n, ip=ip, enc=enc, processed=done,
                                  out_crc=zlib.crc32(out.tobytes())))
        enc, _ = ref.bra_convert(fid, blob, 0x4000, 1)
        dec, _ = ref.bra_convert(fid, enc, 0x4000, 0)
        assert np.array_equal(dec[: len(blob) & ~15], blob[: len(blob) & ~15]) or arch in ("sparc",), arch
    raw = blobs["arm"]
    for dist in (1, 2, 3, 4, 17, 255, 256):
        for n in (0, 1, dist, dist + 1, 999, len(raw)):
            cases.append(dict(kind="delta_enc", blob="arm", dist=dist, n=n, out_crc=zlib.crc32(ref.delta(raw[:n], dist, 1).tobytes())))
    np.savez_compressed(os.path.join(OUT, "bra.npz"), **blobs)
    json.dump(dict(generator="tools/make_golden_bra.py", cases=cases), open(os.path.join(OUT, "bra.json"), "w"), separators=(",", ":"))
    changed = sum(1 for c in cases if c["kind"] == "bra" and c["n"] > 1000 and c["out_crc"] != zlib.crc32(blobs[c["blob"]][:c["n"]].tobytes()))
    print(len(cases), "cases;", changed, "large cases where the converter changed bytes")


if __name__ == "__main__":
    main()
