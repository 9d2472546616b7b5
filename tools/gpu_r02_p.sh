#!/bin/bash
# Round 2, GPU session P: tail policy of dense launches (at most one unit per warp once fewer units than warps are left) A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-p}
NT=$PWD/lzma-sdk-zliblike_b200/lib/variants/notail/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
for n in 8192 6000 10656; do
  run ${T}_l2_${n}_tail   $B --workload lzma2_8g --units $n
  run ${T}_l2_${n}_notail env LZB_LIB=$NT $B --workload lzma2_8g --units $n
done
run ${T}_64k_tail   $B --workload lzma64k
run ${T}_64k_notail env LZB_LIB=$NT $B --workload lzma64k
timeout 900 python -m pytest tests -m gpu -q -x > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
( LZB_GROUP=16 timeout 500 python tools/fuzz_gpu.py --rounds 40 --seed 21 ) > $O/${T}_fuzz_lean16.log 2>&1; echo "fuzz lean G=16 rc=$? $(tail -1 $O/${T}_fuzz_lean16.log)"
