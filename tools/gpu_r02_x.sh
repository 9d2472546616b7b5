#!/bin/bash
# Round 2, GPU session X: share of the fused post pass (x86 BCJ + CRC-64) in the xz4m launch, from a source-level capture of
# the FULL-SIZE launch (2,220 blocks of 4 MiB, one unit per warp, look-ahead form).
set -u
mkdir -p gpurun_out
O=gpurun_out
ARGS="--workload xz4m --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
python bench.py $ARGS > $O/x_xz4m_plain.json 2> $O/x_xz4m_plain.err; echo "plain rc=$?"
timeout 900 ncu --clock-control none --import-source on --section SpeedOfLight --section WarpStateStats --section SchedulerStats \
  --section InstructionStats --section LaunchStats --section SourceCounters -k regex:lzb_decode_units -s 3 -c 1 -f -o $O/x_xz4m \
  python bench.py $ARGS > $O/x_xz4m_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/x_xz4m.ncu-rep --page raw --csv > $O/x_xz4m.raw.csv 2>/dev/null
ncu -i $O/x_xz4m.ncu-rep --page source --csv > $O/x_xz4m.src.csv 2>/dev/null
python tools/ncu_summary.py $O/x_xz4m.raw.csv
python tools/ncu_by_line.py $O/x_xz4m.src.csv lzma-sdk-zliblike_b200/lib/liblzb200.so ILb1ELj32ELb0ELb1E --top 400 > $O/x_xz4m_by_line.txt 2>&1
grep -c . $O/x_xz4m_by_line.txt
rm -f $O/x_xz4m.ncu-rep
