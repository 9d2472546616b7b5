#!/bin/bash
# Round 2, GPU session Y: what the fused post pass (x86 BCJ + CRC-64) costs in the xz4m launch: kernel time with the whole
# pass, without it, and with each half alone (LZB_BENCH_POST, bench.py).
set -u
mkdir -p gpurun_out
O=gpurun_out
for m in all none crc x86; do
  LZB_BENCH_POST=$m timeout 300 python bench.py --workload xz4m --no-cpu --no-e2e --steps 3 --warmup 3 --extras none > $O/y_xz4m_$m.json 2> $O/y_xz4m_$m.err
  echo "$m rc=$? $(python -c "import json;d=json.load(open('$O/y_xz4m_$m.json'));print(round(d['value'],3),round(d['ms_per_step'],2),d['verified'])" 2>&1 | tail -1)"
done
