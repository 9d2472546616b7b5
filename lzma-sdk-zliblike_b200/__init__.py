"""B200-native batched LZMA / LZMA2 / xz decode path behind the LZMA SDK 9.20 C API.

The product is ``lib/liblzb200.so`` (C-ABI declared in ``include/lzb200.h``: the SDK's decode entry
points plus a batched extension), built from ``csrc/`` by ``build()``.  This package is the thin
host-side mirror used by the tests and ``bench.py``: it loads that library with ctypes and calls it.
There is NO CPU fallback: every decode entry point raises if the library or a CUDA device is missing.

Submodules are imported lazily so that ``records`` (plain numpy record layouts) is usable on a
machine without CUDA.
"""
from __future__ import annotations

import importlib

from . import records  # noqa: F401  (no CUDA dependency)

__all__ = ["records", "build", "lib", "sdk", "batch"]


def __getattr__(name):
    if name in ("lib", "sdk", "sharding", "buildlib"):
        return importlib.import_module(f"{__name__}.{name}")
    if name == "build":
        return importlib.import_module(f"{__name__}.buildlib").build
    raise AttributeError(name)
