#!/bin/bash
# Round 2, GPU session A: baselines of the tree as round 1 left it + the LZB_SMALL_LOOP A/B + source-level ncu capture
# of a one-unit-per-warp launch.   gpurun --timeout 1500 -- 'bash tools/gpu_r02_a.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
SMALL=$PWD/lzma-sdk-zliblike_b200/lib/variants/small/liblzb200.so
run() { # name, env..., -- args
  local name=$1; shift
  ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(head -c 300 $O/$name.json | cut -c1-200)"
}
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3"
run a_l2_8192_base   $B --workload lzma2_1m
run a_l2_1024_base   $B --workload lzma2_1m --units 1024
run a_l2_148_base    $B --workload lzma2_1m --units 148
run a_64k_base       $B --workload lzma64k
run a_l2_8192_small  env LZB_LIB=$SMALL $B --workload lzma2_1m
run a_l2_1024_small  env LZB_LIB=$SMALL $B --workload lzma2_1m --units 1024
run a_l2_148_small   env LZB_LIB=$SMALL $B --workload lzma2_1m --units 148
run a_64k_small      env LZB_LIB=$SMALL $B --workload lzma64k
UNITS=1036 bash tools/ncu_few_units.sh
