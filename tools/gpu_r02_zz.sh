#!/bin/bash
# Round 2, last GPU session: the shipped library (lean kernels = session S's, look-ahead / post-only kernels = session Z's
# tile-scan x86 converter): parity suite, smoke, and the two bench lines the change touches.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $O/zz_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/zz_gpu_tests.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/zz_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/zz_smoke.log)"
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
run zz_l2 $B --workload lzma2_8g
run zz_xz4m $B --workload xz4m
