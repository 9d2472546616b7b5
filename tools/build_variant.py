#!/usr/bin/env python
"""Build an A/B copy of liblzb200.so with extra nvcc flags:  python tools/build_variant.py NAME -DFLAG ...
-> lzma-sdk-zliblike_b200/lib/variants/NAME/liblzb200.so; run anything with LZB_LIB=<that path> to use it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lzb200  # noqa: E402

print(lzb200.buildlib.build(variant=sys.argv[1], extra_flags=sys.argv[2:], verbose="-v" in sys.argv))
