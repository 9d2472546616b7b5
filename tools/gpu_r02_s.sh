#!/bin/bash
# Round 2, GPU session S: final validation of the tree -- parity suite (incl. the C host program), smoke, default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-s}
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/${T}_smoke.log)"
( timeout 900 python bench.py ) > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "default bench rc=$?"; tail -8 $O/${T}_bench_default.err
( LZB_GROUP=16 timeout 300 python tools/fuzz_gpu.py --rounds 20 --seed 301 ) > $O/${T}_fuzz_lean16.log 2>&1; echo "fuzz lean rc=$? $(tail -1 $O/${T}_fuzz_lean16.log)"
( timeout 300 python tools/fuzz_gpu.py --rounds 20 --seed 302 ) > $O/${T}_fuzz_ahead.log 2>&1; echo "fuzz ahead rc=$? $(tail -1 $O/${T}_fuzz_ahead.log)"
