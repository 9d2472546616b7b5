#!/bin/bash
# Round 2, GPU session D: parity suite on the rewritten xz framing / BCJ2 merge, dual-form decoder per regime,
# A/B against the round-1 core, first full default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-d}
R1=$PWD/lzma-sdk-zliblike_b200/lib/variants/r1core/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
fi
run ${T}_l2_8192   $B --workload lzma2_8g
run ${T}_l2_8192_r1 env LZB_LIB=$R1 $B --workload lzma2_8g
run ${T}_l2_1024   $B --workload lzma2_8g --units 1024
run ${T}_l2_1024_noahead env LZB_AHEAD_WARPS=0 $B --workload lzma2_8g --units 1024
run ${T}_l2_148    $B --workload lzma2_8g --units 148
run ${T}_l2_2220   $B --workload lzma2_8g --units 2220
run ${T}_l2_2220_ahead env LZB_AHEAD_WARPS=16 $B --workload lzma2_8g --units 2220
run ${T}_l2_1480_ahead env LZB_AHEAD_WARPS=16 $B --workload lzma2_8g --units 1480
run ${T}_l2_1480_lean env LZB_AHEAD_WARPS=0 $B --workload lzma2_8g --units 1480
run ${T}_64k       $B --workload lzma64k
run ${T}_64k_r1    env LZB_LIB=$R1 $B --workload lzma64k
run ${T}_xz4m      $B --workload xz4m
run ${T}_xz4m_r1   env LZB_LIB=$R1 $B --workload xz4m
( timeout 900 python bench.py --steps 2 --warmup 3 ) > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "default bench rc=$?"; tail -12 $O/${T}_bench_default.err; head -c 1200 $O/${T}_bench_default.json
