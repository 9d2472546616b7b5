"""Golden vectors for the 7z folder path: archives written by tools/sz_writer.py (payloads compressed by the
reference's encoders) and what the unmodified reference (SzArEx_Open + SzArEx_Extract per file, through
oracle/ref_harness.c: refh_7z_extract_all) makes of them, pristine and damaged.

    python tools/make_golden_7z.py   ->  tests/golden/sz.json + sz.npz
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from tools import sz_writer as W  # noqa: E402
from tools.corpus import add_x86_callsites, make_corpus  # noqa: E402


def archives(ref):
    c = make_corpus(900000, seed=21)
    code = add_x86_callsites(c[:70000], seed=4).tobytes()
    arm = np.random.default_rng(3).integers(0, 256, 30000, dtype=np.uint8)
    arm[3::28] = 0xEB
    out = {}
    # a: one of everything, plain header
    fa = [W.Folder("lzma", [("a.txt", c[70000:80000].tobytes()), ("dir/b.bin", c[80000:105000].tobytes())]),
          W.Folder("lzma2", [("c.dat", c[110000:150000].tobytes())], file_crc=False, dict_size=1 << 15),
          W.Folder("copy", [("d.raw", c[150000:151000].tobytes())]),
          W.Folder("lzma", [("code.exe", code)], filt="bcj"),
          W.Folder("lzma2", [("arm.bin", arm.tobytes())], filt="arm"),
          W.Folder("lzma", [("nocrc.bin", c[160000:170000].tobytes())], folder_crc=False, file_crc=False)]
    out["a_plain"] = (W.write_archive(ref, fa, empty_files=["empty.txt"], dirs=["dir"]), fa, ["empty.txt"], ["dir"])
    out["a_packed_header"] = (W.write_archive(ref, fa, empty_files=["empty.txt"], dirs=["dir"], encode_header=True), fa, ["empty.txt"], ["dir"])
    # b: many small files in many folders, low entropy to random (BASELINE config 5), packed header
    fb = []
    rng = np.random.default_rng(17)
    for i in range(64):
        n = int(rng.integers(200, 6000))
        frac = i / 63.0
        data = np.where(rng.random(n) < frac, rng.integers(0, 256, n), (np.arange(n) // 7) % 13).astype(np.uint8).tobytes()
        files = [(f"f{i:02d}_{j}.bin", data[j * (n // 3):(j + 1) * (n // 3) if j < 2 else n]) for j in range(3)] if i % 4 == 0 else [(f"f{i:02d}.bin", data)]
        fb.append(W.Folder("lzma2" if i % 3 == 0 else "lzma", files, filt="bcj" if i % 5 == 0 else None, dict_size=1 << 14))
    out["b_many"] = (W.write_archive(ref, fb, encode_header=True), fb, [], [])
    # c: what the reference refuses: Delta as second coder
    fc = [W.Folder("lzma", [("ok.bin", c[200000:205000].tobytes())]), W.Folder("lzma", [("delta.bin", c[210000:215000].tobytes())], filt="delta")]
    out["c_delta_unsupported"] = (W.write_archive(ref, fc), fc, [], [])
    # e: BCJ2 folders (what 7-Zip writes for x86 executables): three LZMA streams + the four-stream merge
    code2 = add_x86_callsites(c[300000:420000], seed=14).tobytes()
    fe = [W.Folder("lzma", [("readme.txt", c[430000:436000].tobytes())]),
          W.Folder("lzma", [("prog.exe", code2[:70000]), ("lib.dll", code2[70000:])], filt="bcj2"),
          W.Folder("lzma", [("tiny.exe", bytes([0x90, 0xE8, 1, 2, 3, 4, 0xC3, 0xE9]))], filt="bcj2"),
          W.Folder("lzma2", [("tail.bin", c[440000:450000].tobytes())])]
    out["e_bcj2"] = (W.write_archive(ref, fe, encode_header=True), fe, [], [])
    # d: only empty entries
    out["d_empty_only"] = (W.write_archive(ref, [], empty_files=["e1", "e2"], dirs=["d1"]), [], ["e1", "e2"], ["d1"])
    return out


def summarize(d):
    files = [dict(res=f["res"], size=f["size"], folder=f["folder"] if f["folder"] != 0xFFFFFFFF else -1, is_dir=f["is_dir"],
                  has_stream=f["has_stream"], crc_defined=f["crc_defined"], name=f["name"],
                  data_crc=int(zlib.crc32(d["out"][f["off"]:f["off"] + f["size"]].tobytes())) if f["res"] == 0 and f["has_stream"] else None)
             for f in d["files"]]
    return dict(open_res=d["open_res"], num_files=d["num_files"], num_folders=d["num_folders"], files=files)


def mutations(name, arc, rng):
    n = len(arc)
    ms = [("xor", 0, 1), ("xor", 6, 1), ("xor", 9, 1), ("xor", 13, 0x40), ("xor", 21, 1), ("xor", 29, 1), ("cut", n - 1), ("cut", n // 2), ("cut", 31),
          ("xor", n - 3, 0x10), ("xor", n - 20, 0x04), ("xor", 40, 0x01), ("xor", 33, 0x80), ("xor", n // 3, 0x20), ("xor", n // 2, 0x02)]
    for _ in range(14):
        ms.append(("xor", int(rng.integers(32, n)), 1 << int(rng.integers(0, 8))))
    return ms


def mutate(b, m):
    if m is None:
        return b
    if m[0] == "cut":
        return b[:m[1]]
    a = bytearray(b)
    a[m[1]] ^= m[2]
    return bytes(a)


def main():
    from oracle import refbind
    ref = refbind.ref()
    rng = np.random.default_rng(99)
    blobs, cases = {}, []
    for name, (arc, folders, empties, dirs) in archives(ref).items():
        blobs[name] = np.frombuffer(arc, dtype=np.uint8)
        d = ref.sz_extract_all(arc)
        s = summarize(d)
        # the writer and the reference agree on the pristine archive
        exp = W.expected_listing(folders, empties, dirs)
        assert s["open_res"] == 0 and s["num_files"] == len(exp), (name, s["open_res"])
        for f, (nm, data) in zip(s["files"], exp):
            assert f["name"] == nm, (name, nm)
            if data is not None and f["res"] == 0:
                assert f["data_crc"] is None or f["data_crc"] == zlib.crc32(data), (name, nm)
        cases.append(dict(blob=name, mut=None, **s))
        for m in mutations(name, arc, rng):
            cases.append(dict(blob=name, mut=list(m), **summarize(ref.sz_extract_all(mutate(arc, m)))))
    np.savez_compressed(os.path.join(OUT, "sz.npz"), **blobs)
    json.dump(dict(generator="tools/make_golden_7z.py", cases=cases), open(os.path.join(OUT, "sz.json"), "w"), separators=(",", ":"))
    import collections
    print(len(cases), "cases; open_res:", dict(collections.Counter(c["open_res"] for c in cases)),
          "file res:", dict(collections.Counter(f["res"] for c in cases for f in c["files"])))


if __name__ == "__main__":
    main()
