#!/bin/bash
# Round 2, GPU session N: xz / 7z parity after the last text edits of the framing code; per-GPU unit counts between one
# unit per warp and full residency: lean (two per warp) vs look-ahead forced to one unit per warp, persistent.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-n}
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
timeout 900 python -m pytest tests/test_xz_stream.py tests/test_7z.py tests/test_gpu_parity.py -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
for n in 3072 4096 5328 8192; do
  run ${T}_l2_${n}_auto $B --workload lzma2_8g --units $n
  run ${T}_l2_${n}_g32  env LZB_GROUP=32 $B --workload lzma2_8g --units $n
done
run ${T}_64k_g32  env LZB_GROUP=32 $B --workload lzma64k
