#!/bin/bash
# Round 2, kernel log M: -DLZB_LEAN_EARLY_LOAD (the child's probability requested before the node's is stored back) against
# the shipped library, dense workload, same box.
set -u
O=gpurun_out
V=$PWD/lzma-sdk-zliblike_b200/lib/variants/early/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 --extras none --workload lzma2_8g"
run m_l2_early env LZB_LIB=$V $B
run m_l2_shipped $B
