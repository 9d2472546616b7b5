#!/bin/bash
# Round 2, GPU session T: evidence of the SHIPPED kernels -- ncu launch list of exactly the default `python bench.py`,
# `--set full` captures of the lean form (dense launch) and the look-ahead form (one unit per warp) of the final tree.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-t}
# 1. full captures at replayable sizes (after the same commands exited 0 without ncu)
D_ARGS="--workload lzma64k --units 16384 --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
A_ARGS="--workload lzma2_8g --units 1024 --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
python bench.py $D_ARGS > $O/${T}_dense_plain.json 2> $O/${T}_dense_plain.err; echo "dense plain rc=$?"
python bench.py $A_ARGS > $O/${T}_ahead_plain.json 2> $O/${T}_ahead_plain.err; echo "ahead plain rc=$?"
for v in dense ahead; do
  A="$D_ARGS"; [ $v = ahead ] && A="$A_ARGS"
  timeout 600 ncu --clock-control none --import-source on --set full -k regex:lzb_decode_units -s 3 -c 1 -f -o $O/${T}_$v \
    python bench.py $A > $O/${T}_${v}_ncu.log 2>&1; echo "ncu $v rc=$?"
  ncu -i $O/${T}_$v.ncu-rep --page raw --csv > $O/${T}_$v.raw.csv 2>/dev/null
  ncu -i $O/${T}_$v.ncu-rep --page source --csv > $O/${T}_$v.src.csv 2>/dev/null
  python tools/ncu_summary.py $O/${T}_$v.raw.csv > $O/${T}_$v.summary.txt 2>&1
  cat $O/${T}_$v.summary.txt
done
# 2. launch list of the default bench, the exact command the driver runs (extras included)
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "default bench rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches_default.csv \
  python bench.py > $O/${T}_launches_default.log 2>&1; echo "launch list rc=$?"
grep -c lzb_ $O/${T}_launches_default.csv
rm -f $O/*.ncu-rep.tmp; ls -la $O/${T}_*
