#!/bin/bash
# Source-level ncu capture of a ONE-UNIT-PER-WARP launch (G = 32: fewer units than warps), at a size ncu can replay.
#
#   gpurun --timeout 600 -- 'bash tools/ncu_few_units.sh'      (one GPU; never under torchrun)
#
# The full-size xz4m launch (1.7 s, 14 GB of buffers saved and restored per replay pass) does not finish one pass
# in six minutes of box time (profiles/r01_j_few_units.md).  1,036 units of 1 MiB = 7 per SM keep the mode
# (matched-literal trees in shared memory, spread over all SMs) with a 0.35 s kernel and 1.5 GB of buffers.
# Writes gpurun_out/few_units.{json,ncu-rep,raw.csv,src.csv}; copy what is to be judged into profiles/.
set -u
UNITS=${UNITS:-1036}
ARGS="--workload lzma2_1m --units $UNITS --no-cpu --no-e2e --steps 2 --warmup 3"
mkdir -p gpurun_out
# 1. the same command must exit 0 without ncu first
timeout 120 python bench.py $ARGS > gpurun_out/few_units.json 2> gpurun_out/few_units.err || { echo "plain run failed"; exit 1; }
# 2. one launch, sections that carry the issue-slot / stall breakdown and the per-line samples
timeout 400 ncu --clock-control none --import-source on \
  --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section InstructionStats \
  --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis --section SourceCounters \
  -k regex:lzb_decode_units -s 3 -c 1 -f -o gpurun_out/few_units python bench.py $ARGS > gpurun_out/few_units_ncu.log 2>&1
tail -3 gpurun_out/few_units_ncu.log
[ -f gpurun_out/few_units.ncu-rep ] || exit 2
ncu -i gpurun_out/few_units.ncu-rep --page raw --csv > gpurun_out/few_units.raw.csv 2>/dev/null
ncu -i gpurun_out/few_units.ncu-rep --page source --csv > gpurun_out/few_units.src.csv 2>/dev/null
ls -la gpurun_out/few_units.*
