#!/bin/bash
# Round 2, GPU session K: look-ahead walk, bit step shaped for a short chain (v2, default) vs the first version (v1).
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-k}
V1=$PWD/lzma-sdk-zliblike_b200/lib/variants/aheadv1/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
for n in 148 1024 2048; do
  run ${T}_l2_${n}_v2 $B --workload lzma2_8g --units $n
  run ${T}_l2_${n}_v1 env LZB_LIB=$V1 $B --workload lzma2_8g --units $n
done
run ${T}_xz4m_v2 $B --workload xz4m
run ${T}_xz4m_v1 env LZB_LIB=$V1 $B --workload xz4m
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or golden or stream or 7z" > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
