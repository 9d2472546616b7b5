#!/usr/bin/env python
"""Per-kernel comparison of the SASS of two builds of liblzb200.so (labels and branch-target comments aside):

    python tools/sass_compare.py OLD.so NEW.so

Used to show that a change of cold code left a kernel's instructions untouched (profiles/r02_z_x86_tiles.md): the hot
loop's register allocation moves with whatever is inlined next to it, so "same source in the hot path" is not enough."""
import os
import re
import subprocess
import sys
import tempfile


def kernels(so):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if "lzb_kernels" in f][0]
    txt = subprocess.run(["nvdisasm", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    starts = [i for i, l in enumerate(txt) if l.startswith(".text.")]
    out = {}
    for k, s in enumerate(starts):
        e = starts[k + 1] if k + 1 < len(starts) else len(txt)
        ins = []
        for l in txt[s:e]:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                t = re.sub(r"`\(\.L_x_\d+\)", "L", m.group(2))
                t = re.sub(r"\(\*\"BRANCH_TARGETS[^)]*\)", "", t)
                t = re.sub(r"`\(\$[^)]*\)", "CALLEE", t)          # names of out-of-line device functions
                ins.append(t.strip())
        out[txt[s].strip()[6:].rstrip(":")] = ins
    return out


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    for k in sorted(set(a) | set(b)):
        x, y = a.get(k), b.get(k)
        if x is None or y is None:
            verdict = "only in " + ("OLD" if y is None else "NEW")
        elif x == y:
            verdict = f"identical ({len(x)} instructions)"
        else:
            verdict = f"DIFFERENT ({len(x)} -> {len(y)} instructions, {sum(p != q for p, q in zip(x, y))} differ in place)"
        print(f"{verdict:60s} {k[:100]}")


if __name__ == "__main__":
    main()
