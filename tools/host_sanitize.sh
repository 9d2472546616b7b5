#!/bin/bash
# The HOST side of liblzb200.so (lzb_host.cu, lzb_state.cu, lzb_containers.cu, lzb_xzstream.cu, lzb_7z.cu) built with
# AddressSanitizer + UBSan (device code unchanged) and driven by the test suite:
#     tools/host_sanitize.sh            CPU tests (container walks, hostile 7z databases, ABI, sharding)
#     tools/host_sanitize.sh gpu        on a GPU box: the whole `-m gpu` suite through the sanitized host code
# (test_c_host links a C program against the library and cannot link the sanitized one: left out.)
set -u
cd "$(dirname "$0")/.."
SANF="-Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-omit-frame-pointer -Xcompiler -fno-sanitize-recover=undefined"
LIB=$PWD/lzma-sdk-zliblike_b200/lib/variants/san/liblzb200.so
[ "${1:-}" = "gpu" ] && [ -f "$LIB" ] || python tools/build_variant.py san $SANF | tail -1
ASAN=$(gcc -print-file-name=libasan.so)
UBSAN=$(gcc -print-file-name=libubsan.so)
export LZB_LIB=$LIB LD_PRELOAD="$ASAN:$UBSAN" ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
if [ "${1:-}" = "gpu" ]; then
  shift
  python -m pytest tests -q -m gpu --deselect tests/test_c_host.py "$@"
else
  python -m pytest tests/test_abi_and_host.py tests/test_sharding_gloo.py tests/test_7z.py tests/test_xz_stream.py -q -m "not gpu" "$@"
fi
