#!/bin/bash
# Round 2, GPU session Z: x86 BCJ as a tile scan (coalesced 16-byte loads, candidates walked in order) instead of one
# byte stream per lane: GPU parity suite, then A/B against the library of session S (variants/old) and the inlined build.
set -u
mkdir -p gpurun_out
O=gpurun_out
OLD=$PWD/lzma-sdk-zliblike_b200/lib/variants/old/liblzb200.so
INL=$PWD/lzma-sdk-zliblike_b200/lib/variants/x86inl/liblzb200.so
timeout 900 python -m pytest tests -x -q -m gpu > $O/z_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/z_gpu_tests.log)"
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python -c "import json;d=json.load(open('$O/$name.json'));print(round(d['value'],3),round(d['ms_per_step'],1),d['verified'])" 2>&1 | tail -1)"; }
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --extras none"
run z_xz4m_new $B --workload xz4m
run z_xz4m_inl env LZB_LIB=$INL $B --workload xz4m
run z_l2_new $B --workload lzma2_8g
run z_l2_old env LZB_LIB=$OLD $B --workload lzma2_8g
run z_l2_inl env LZB_LIB=$INL $B --workload lzma2_8g
run z_64k_new $B --workload lzma64k
run z_64k_old env LZB_LIB=$OLD $B --workload lzma64k
timeout 300 python tools/bench_unbatched.py > $O/z_unbatched.json 2> $O/z_unbatched.err; echo "unbatched rc=$?"
python -c "
import json;d=json.load(open('$O/z_unbatched.json'));print({k:v for k,v in d.items() if 'x86' in k.lower()})"
