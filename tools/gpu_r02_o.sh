#!/bin/bash
# Round 2, GPU session O: differential fuzz against the live reference in all three decoder configurations (look-ahead
# form; lean form with two units per warp; lean form with one unit per warp), multi-unit LZMA2 fuzz, and the new
# geometry between one wave of warps and the dense regime.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-o}
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
run ${T}_l2_4096 $B --workload lzma2_8g --units 4096
run ${T}_l2_3072 $B --workload lzma2_8g --units 3072
( timeout 500 python tools/fuzz_gpu.py --rounds 30 --seed 11 ) > $O/${T}_fuzz_ahead.log 2>&1; echo "fuzz ahead rc=$? $(tail -1 $O/${T}_fuzz_ahead.log)"
( LZB_GROUP=16 timeout 500 python tools/fuzz_gpu.py --rounds 30 --seed 12 ) > $O/${T}_fuzz_lean16.log 2>&1; echo "fuzz lean G=16 rc=$? $(tail -1 $O/${T}_fuzz_lean16.log)"
( LZB_AHEAD_WARPS=0 timeout 500 python tools/fuzz_gpu.py --rounds 20 --seed 13 ) > $O/${T}_fuzz_lean32.log 2>&1; echo "fuzz lean G=32 rc=$? $(tail -1 $O/${T}_fuzz_lean32.log)"
( timeout 500 python tools/fuzz_gpu.py --lzma2 --rounds 25 --seed 14 ) > $O/${T}_fuzz_l2_ahead.log 2>&1; echo "fuzz lzma2 ahead rc=$? $(tail -1 $O/${T}_fuzz_l2_ahead.log)"
( LZB_GROUP=16 timeout 500 python tools/fuzz_gpu.py --lzma2 --rounds 25 --seed 15 ) > $O/${T}_fuzz_l2_lean.log 2>&1; echo "fuzz lzma2 lean rc=$? $(tail -1 $O/${T}_fuzz_l2_lean.log)"
