#!/bin/bash
# Round 2, GPU session B: the predicated-normalisation kernel -- parity suite, benches in all regimes, ncu captures.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-b}
run() { local name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(head -c 200 $O/$name.json)"; }
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
fi
run ${T}_l2_8192   $B --workload lzma2_1m
run ${T}_l2_1024   $B --workload lzma2_1m --units 1024
run ${T}_l2_148    $B --workload lzma2_1m --units 148
run ${T}_64k       $B --workload lzma64k
run ${T}_xz4m      $B --workload xz4m
if [ "${SKIP_NCU:-0}" != "1" ]; then
  UNITS=1036 bash tools/ncu_few_units.sh
  for f in few_units.json few_units.ncu-rep few_units.raw.csv few_units.src.csv; do mv $O/$f $O/${T}_$f; done
  # throughput regime: 16,384 x 64 KiB units (two per warp, 36 per SM)
  ARGS="--workload lzma64k --units 16384 --no-cpu --no-e2e --steps 2 --warmup 3"
  timeout 600 ncu --clock-control none --import-source on --set full -k regex:lzb_decode_units -s 3 -c 1 -f -o $O/${T}_64k16k python bench.py $ARGS > $O/${T}_64k16k_ncu.log 2>&1
  ncu -i $O/${T}_64k16k.ncu-rep --page raw --csv > $O/${T}_64k16k.raw.csv 2>/dev/null
  ncu -i $O/${T}_64k16k.ncu-rep --page source --csv > $O/${T}_64k16k.src.csv 2>/dev/null
  ls -la $O/${T}_*
fi
