#!/bin/bash
# Round 2: look-ahead kernels without any filter (the per-GPU shares of config 3 at 8 and 4 GPUs), shipped library vs the
# library of session S -- did the out-of-line x86 tile scan next to the look-ahead decoder move its hot loop?
set -u
O=gpurun_out
OLD=$PWD/lzma-sdk-zliblike_b200/lib/variants/old/liblzb200.so
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none --workload lzma2_8g"
run zz3_1024_new $B --units 1024
run zz3_1024_old env LZB_LIB=$OLD $B --units 1024
run zz3_2048_new $B --units 2048
run zz3_2048_old env LZB_LIB=$OLD $B --units 2048
