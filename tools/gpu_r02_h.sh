#!/bin/bash
# Round 2, GPU session H: parity suite after the stateful-staging fix; dense-launch A/B: post pass out of line vs inline
# vs the round-1 core.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-h}
R1=$PWD/lzma-sdk-zliblike_b200/lib/variants/r1core/liblzb200.so
PI=$PWD/lzma-sdk-zliblike_b200/lib/variants/postinline/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
for w in lzma2_8g lzma64k; do
  run ${T}_${w}_cold   $B --workload $w
  run ${T}_${w}_inline env LZB_LIB=$PI $B --workload $w
  run ${T}_${w}_r1     env LZB_LIB=$R1 $B --workload $w
done
run ${T}_l2_1024   $B --workload lzma2_8g --units 1024
run ${T}_xz4m      $B --workload xz4m
