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
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
# extra nvcc flags for A/B builds of kernel experiments (forced rebuild needed); the shipped library is built without any
NVCC_FLAGS += os.environ.get("LZB_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblzb200.so cannot be built (there is no CPU fallback)")


def _source_digest(extra_flags=()) -> str:
    """What the library is made of: every file of csrc/, the public header and the compiler flags."""
    import hashlib
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), "include", "lzb200.h")
    for path in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [inc]:
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    h.update(" ".join([*NVCC_FLAGS, *extra_flags]).encode())
    return h.hexdigest()


def _stale() -> bool:
    """By content, not by time stamps: a copy of the tree (the GPU box gets one) keeps a library that matches its
    sources whatever the copy did to the mtimes."""
    stamp = SO + ".src"
    if not os.path.exists(SO) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != _source_digest()


def build(force: bool = False, verbose: bool = False, variant: str | None = None, extra_flags=()) -> str:
    """Compile (if needed) and return the path of liblzb200.so.

    variant: build an A/B copy of the library into lib/variants/<variant>/ with extra nvcc flags (kernel
    experiments; selected at run time with LZB_LIB=<path>).  The shipped library is built without any.

    Safe against several processes calling it at once (one rank per GPU under torchrun): the build happens under an
    exclusive file lock, into a private directory, and the finished library is moved into place atomically -- a
    process never sees a half-written file, and whoever comes second finds the library fresh."""
    if variant is None and os.environ.get("LZB_LIB"):
        return os.environ["LZB_LIB"]                    # an A/B build chosen for this process (tools/build_variant.py)
    if variant is None and not force and not _stale():
        return SO
    libdir = LIBDIR if variant is None else os.path.join(LIBDIR, "variants", variant)
    so = os.path.join(libdir, "liblzb200.so")
    os.makedirs(libdir, exist_ok=True)
    import fcntl
    import tempfile
    with open(os.path.join(libdir, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if variant is None and not force and not _stale():
            return so                                   # another process built it while this one waited
        nvcc = _nvcc()
        work = tempfile.mkdtemp(prefix=".build-", dir=libdir)
        try:
            objs, cmds = [], []
            for src in SOURCES:
                obj = os.path.join(work, src.replace(".cu", ".o"))
                cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                    print(" ".join(cmd))
                cmds.append(cmd)
                objs.append(obj)
            procs = [subprocess.Popen(c) for c in cmds]      # the translation units are independent: side by side
            for c, p in zip(cmds, procs):
                if p.wait() != 0:
                    raise subprocess.CalledProcessError(p.returncode, c)
            tmp_so = os.path.join(work, "liblzb200.so")
            cmd = [nvcc, "-shared", "-o", tmp_so, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                   "-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
            subprocess.check_call(cmd)
            os.replace(tmp_so, so)
            with open(so + ".src", "w") as f:
                f.write(_source_digest(extra_flags) + "\n")
        finally:
            shutil.rmtree(work, ignore_errors=True)
    return so


if __name__ == "__main__":
    print(build(force=True, verbose=True))
