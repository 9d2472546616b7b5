#!/bin/bash
# Round 2, GPU session R: long differential fuzz of the final tree in all decoder configurations + xz / 7z / streaming suites.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r}
( timeout 900 python tools/fuzz_gpu.py --rounds 200 --seed 101 ) > $O/${T}_fuzz_ahead.log 2>&1; echo "fuzz ahead rc=$? $(tail -1 $O/${T}_fuzz_ahead.log)"
( LZB_GROUP=16 timeout 900 python tools/fuzz_gpu.py --rounds 200 --seed 102 ) > $O/${T}_fuzz_lean16.log 2>&1; echo "fuzz lean G=16 rc=$? $(tail -1 $O/${T}_fuzz_lean16.log)"
( LZB_GROUP=8 timeout 600 python tools/fuzz_gpu.py --rounds 60 --seed 103 ) > $O/${T}_fuzz_lean8.log 2>&1; echo "fuzz lean G=8 rc=$? $(tail -1 $O/${T}_fuzz_lean8.log)"
( LZB_AHEAD_WARPS=0 timeout 600 python tools/fuzz_gpu.py --rounds 100 --seed 104 ) > $O/${T}_fuzz_lean32.log 2>&1; echo "fuzz lean G=32 rc=$? $(tail -1 $O/${T}_fuzz_lean32.log)"
( timeout 900 python tools/fuzz_gpu.py --lzma2 --rounds 150 --seed 105 ) > $O/${T}_fuzz_l2_ahead.log 2>&1; echo "fuzz lzma2 ahead rc=$? $(tail -1 $O/${T}_fuzz_l2_ahead.log)"
( LZB_GROUP=16 timeout 900 python tools/fuzz_gpu.py --lzma2 --rounds 150 --seed 106 ) > $O/${T}_fuzz_l2_lean.log 2>&1; echo "fuzz lzma2 lean rc=$? $(tail -1 $O/${T}_fuzz_l2_lean.log)"
timeout 1200 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
