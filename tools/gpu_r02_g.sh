#!/bin/bash
# Round 2, GPU session G: full parity suite, lean form (round-1 text in its own include) vs round-1 core, look-ahead
# launches, the default bench line, unbatched table.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-g}
R1=$PWD/lzma-sdk-zliblike_b200/lib/variants/r1core/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --extras none"
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
run ${T}_l2_8192   $B --workload lzma2_8g
run ${T}_l2_8192_r1 env LZB_LIB=$R1 $B --workload lzma2_8g
run ${T}_64k       $B --workload lzma64k
run ${T}_64k_r1    env LZB_LIB=$R1 $B --workload lzma64k
run ${T}_l2_1024   $B --workload lzma2_8g --units 1024
run ${T}_l2_148    $B --workload lzma2_8g --units 148
run ${T}_xz4m      $B --workload xz4m
( timeout 300 python tools/bench_unbatched.py ) > $O/${T}_unbatched.json 2> $O/${T}_unbatched.err; echo "unbatched rc=$?"; cut -c1-600 $O/${T}_unbatched.json
( timeout 900 python bench.py ) > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "default bench rc=$?"; tail -8 $O/${T}_bench_default.err; head -c 600 $O/${T}_bench_default.json
