#!/bin/bash
# Round 2, GPU session E: profiles -- look-ahead form at 7 units/SM (source level), lean form now vs round-1 core
# (instruction counts), DRAM traffic of the full-size launches for profiles/traffic.json.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-e}
R1=$PWD/lzma-sdk-zliblike_b200/lib/variants/r1core/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
run ${T}_xz4m_ahead16 env LZB_AHEAD_WARPS=16 $B --workload xz4m
run ${T}_l2_2368_ahead16 env LZB_AHEAD_WARPS=16 $B --workload lzma2_8g --units 2368
run ${T}_l2_2368_lean env LZB_AHEAD_WARPS=0 $B --workload lzma2_8g --units 2368
# 1. look-ahead form, one unit per warp, 7 per SM
ARGS="--workload lzma2_8g --units 1036 --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
timeout 500 ncu --clock-control none --import-source on --section SpeedOfLight --section WarpStateStats --section SchedulerStats \
  --section InstructionStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section SourceCounters \
  -k regex:lzb_decode_units -s 3 -c 1 -f -o $O/${T}_few python bench.py $ARGS > $O/${T}_few_ncu.log 2>&1
ncu -i $O/${T}_few.ncu-rep --page raw --csv > $O/${T}_few.raw.csv 2>/dev/null
ncu -i $O/${T}_few.ncu-rep --page source --csv > $O/${T}_few.src.csv 2>/dev/null
# 2. lean form vs round-1 core, 16,384 x 64 KiB
ARGS="--workload lzma64k --units 16384 --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
for v in cur r1; do
  L=""; [ $v = r1 ] && L="LZB_LIB=$R1"
  env $L timeout 500 ncu --clock-control none --import-source on --set full -k regex:lzb_decode_units -s 3 -c 1 -f -o $O/${T}_64k_$v python bench.py $ARGS > $O/${T}_64k_${v}_ncu.log 2>&1
  ncu -i $O/${T}_64k_$v.ncu-rep --page raw --csv > $O/${T}_64k_$v.raw.csv 2>/dev/null
  ncu -i $O/${T}_64k_$v.ncu-rep --page source --csv > $O/${T}_64k_$v.src.csv 2>/dev/null
done
# 3. DRAM traffic of the full-size launches (one pass, no replay of other metrics)
for w in lzma2_8g xz4m lzma64k; do
  timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:lzb_decode_units -s 3 -c 1 --csv \
    --log-file $O/${T}_dram_$w.csv python bench.py --workload $w --no-cpu --no-e2e --steps 2 --warmup 3 --extras none > $O/${T}_dram_$w.log 2>&1
  tail -4 $O/${T}_dram_$w.csv
done
rm -f $O/*.ncu-rep.tmp; ls -la $O/${T}_*
