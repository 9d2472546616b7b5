#!/bin/bash
# Round 2, GPU session M: validation of the tree as it stands -- parity suite, smoke, default bench line, unbatched table.
set -u
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-m}
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "pytest rc=$? $(tail -1 $O/${T}_gpu_tests.log)"
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/${T}_smoke.log)"
( timeout 300 python tools/bench_unbatched.py ) > $O/${T}_unbatched.json 2> $O/${T}_unbatched.err; echo "unbatched rc=$?"; python -c "import json;d=json.load(open('$O/${T}_unbatched.json'));print(d['x86_Convert_1MiB'], d['LzmaDec_DecodeToBuf_16_bytes_per_call'], d['speedup_64_callers_over_1'])"
( timeout 900 python bench.py ) > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "default bench rc=$?"; tail -8 $O/${T}_bench_default.err
( timeout 600 python bench.py --impl reference ) > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference arm rc=$?"; head -c 300 $O/${T}_bench_reference.json
