#!/usr/bin/env python
"""Per-source-line table of an ncu capture:  joins `ncu -i X.ncu-rep --page source --csv` (SASS rows with executed
instructions and stall samples) with `nvdisasm -gi` of the same kernel from the shipped cubin (source line of every
SASS instruction, inlining chain included).

  python tools/ncu_by_line.py gpurun_out/few_units.src.csv lzma-sdk-zliblike_b200/lib/liblzb200.so 'ILb1ELj32ELb0E' [--top 40]

Prints: share of executed warp instructions and of stall samples per innermost source line, and per "site" (the
outermost frame inside decode_symbol / step / advance), plus the top stall reasons of each.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def disasm_lines(so, tag):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if "lzb_kernels" in f][0]
    txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    start = [i for i, l in enumerate(txt) if l.startswith(".text.") and "lzb_decode_units_kernel" + tag in l]
    assert start, "kernel not found"
    out, chain = [], []
    fresh = True
    for l in txt[start[0] + 1:]:
        if l.startswith(".text.") or l.startswith("//--------------------- .text"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), tuple(chain)))
            fresh = True
    return out


def main():
    src_csv, so, tag = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    body = rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    dis = disasm_lines(so, tag)
    assert len(dis) == len(body), (len(dis), len(body))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_i = sum(int(r[col["Instructions Executed"]]) for r in body)
    tot_s = sum(int(r[col["# Samples"]]) for r in body)
    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    by_site = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for (off, ins, chain), r in zip(dis, body):
        n_i, n_s = int(r[col["Instructions Executed"]]), int(r[col["# Samples"]])
        inner = chain[0] if chain else ("?", 0)
        # site = the frame just below the kernel body (the last lzb_core.cuh frame) -> what part of the symbol this is
        core = [c for c in chain if c[0] == "lzb_core.cuh"]
        site = core[-1] if core else (chain[-1] if chain else ("?", 0))
        for key, tab in ((inner, by_line), (site, by_site)):
            tab[key][0] += n_i
            tab[key][1] += n_s
            for h in stall_cols:
                v = int(r[col[h]])
                if v:
                    tab[key][2][h[6:]] += v
    print(f"total: {tot_i} warp instructions executed, {tot_s} stall samples, {len(dis)} SASS instructions")
    for name, tab in (("site (outermost lzb_core.cuh frame)", by_site), ("innermost source line", by_line)):
        print(f"\n== by {name}: file:line  instr%  samples%  cycles/instr(rel)  top stalls")
        for key, (n_i, n_s, st) in sorted(tab.items(), key=lambda kv: -kv[1][1])[:top]:
            tops = ", ".join(f"{k} {v * 100 / max(n_s, 1):.0f}%" for k, v in st.most_common(3))
            print(f"{key[0]}:{key[1]:<5d} {n_i * 100 / tot_i:6.2f} {n_s * 100 / tot_s:6.2f}  {(n_s / tot_s) / max(n_i / tot_i, 1e-9):5.2f}  {tops}")


if __name__ == "__main__":
    main()
