#!/bin/bash
# Round 2, GPU session W: the whole GPU suite through the ASan + UBSan build of the host code (tools/host_sanitize.sh gpu),
# then the 7z tests with the shipped library (its two memcpy guards are the only change since session S).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 bash tools/host_sanitize.sh gpu -x > $O/w_host_sanitize_gpu.log 2>&1
echo "sanitized gpu suite rc=$? $(grep -E 'passed|failed' $O/w_host_sanitize_gpu.log | tail -1)"
grep -n "runtime error\|ERROR: AddressSanitizer\|SUMMARY" $O/w_host_sanitize_gpu.log | head -10 | cut -c1-300
timeout 600 python -m pytest tests/test_7z.py tests/test_c_host.py -x -q -m gpu > $O/w_7z_shipped.log 2>&1
echo "shipped 7z + c_host rc=$? $(tail -1 $O/w_7z_shipped.log)"
