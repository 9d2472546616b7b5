#!/usr/bin/env python
"""Key numbers of an `ncu --page raw --csv` export:  python tools/ncu_summary.py X.raw.csv [out_bytes]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
out_bytes = float(sys.argv[2]) if len(sys.argv) > 2 else None


def f(k):
    try:
        return float(m[k].replace(",", ""))
    except Exception:
        return float("nan")


print("kernel", m.get("Kernel Name", "")[:60], "grid", m.get("launch__grid_size"), "block", m.get("launch__block_size"), "regs", m.get("launch__registers_per_thread"))
print("duration ms", f("gpu__time_duration.sum"))
inst = f("smsp__inst_executed.sum")
print("warp instr", inst, "per out byte", inst / out_bytes if out_bytes else None)
print("threads/inst", f("smsp__thread_inst_executed_per_inst_executed.ratio"))
print("ipc (sm)", f("sm__inst_executed.avg.per_cycle_active"), "issue active %", f("smsp__issue_active.avg.pct_of_peak_sustained_active"))
print("eligible warps/sched", f("smsp__warps_eligible.avg.per_cycle_active"), "active warps/sched", f("smsp__warps_active.avg.per_cycle_active"))
print("cycles/inst per warp", f("smsp__average_warp_latency_per_inst_issued.ratio"))
st = {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: f(k) for k in hdr
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k}
print("stalls/issue:", ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v >= 0.02))
print("L1 hit %", f("l1tex__t_sector_hit_rate.pct"), "L2 hit %", f("lts__t_sector_hit_rate.pct"))
print("dram read GB", f("dram__bytes_read.sum") / (1e9 if m.get("dram__bytes_read.sum") else 1), "write", f("dram__bytes_write.sum"), "(units:", rows[1][hdr.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in hdr else "", ")")
