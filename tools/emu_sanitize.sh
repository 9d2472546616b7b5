#!/bin/bash
# The device decode logic (csrc/lzb_core.cuh, lzb_core_lean.inc, lzb_post.cuh built for the host with -DLZB_EMU) under
# AddressSanitizer + UBSan, driven by the emulation tests (golden vectors, ring-wrap and streaming cases, both forms of
# the decoder).  CPU only.  compute-sanitizer is not available on the GPU pool; this checks the same index arithmetic.
set -u
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
UBSAN=$(gcc -print-file-name=libubsan.so)
LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  LZB_EMU_SANITIZE=1 python -m pytest tests/test_golden_cpu.py tests/test_streaming.py tests/test_xz_stream.py -q -m "not gpu" "$@"
