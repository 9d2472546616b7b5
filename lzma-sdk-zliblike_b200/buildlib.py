"""Build lib/liblzb200.so from csrc/ with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "liblzb200.so")
SOURCES = ["lzb_kernels.cu", "lzb_host.cu", "lzb_containers.cu", "lzb_state.cu", "lzb_xzstream.cu", "lzb_7z.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--use_fast_math"]
# extra nvcc flags for A/B builds of kernel experiments (forced rebuild needed); the shipped library is built without any
NVCC_FLAGS += os.environ.get("LZB_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblzb200.so cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    inc = os.path.join(os.path.dirname(HERE), "include", "lzb200.h")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [inc]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if needed) and return the path of liblzb200.so."""
    if not force and not _stale():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
