#!/bin/bash
# Round 2, GPU session J: parity suite; dense launches after the cold GF(2) loops were rolled (kernel 12,400 -> 8,904 SASS
# instructions) vs the round-1 core; unbatched table incl. the resumable-object call cost; ncu launch list of the default bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-j}
R1=$PWD/lzma-sdk-zliblike_b200/lib/variants/r1core/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
for w in lzma2_8g lzma64k; do
  run ${T}_${w}      $B --workload $w
  run ${T}_${w}_r1   env LZB_LIB=$R1 $B --workload $w
done
run ${T}_l2_1024   $B --workload lzma2_8g --units 1024
run ${T}_l2_2048   $B --workload lzma2_8g --units 2048
run ${T}_l2_4096   $B --workload lzma2_8g --units 4096
run ${T}_xz4m      $B --workload xz4m
( timeout 300 python tools/bench_unbatched.py ) > $O/${T}_unbatched.json 2> $O/${T}_unbatched.err; echo "unbatched rc=$?"; python -c "import json;d=json.load(open('$O/${T}_unbatched.json'));print(d['LzmaDec_DecodeToBuf_16_bytes_per_call'], d['speedup_64_callers_over_1'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --extras weak_lzma64k,xz4m > $O/${T}_launches.log 2>&1; echo "launch list rc=$?"; grep -c lzb_ $O/${T}_launches.csv
