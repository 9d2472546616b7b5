"""pytest configuration.

Markers
-------
gpu     : needs a real B200 (the parity tests proper; they call the product through its C-ABI)
needref : needs oracle/_ref/libref.so, i.e. the unmodified reference compiled from /root/reference.
          Present in the build container and (because oracle/_ref/ travels with gpurun) on the GPU
          box; skipped elsewhere -- the committed tests/golden/ vectors cover that case.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "needref: test needs oracle/_ref/libref.so (compiled reference)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    from oracle import refbind
    have_ref = refbind.have_ref()
    have_gpu = None
    for it in items:
        if "needref" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="oracle/_ref/libref.so not built (no /root/reference here)"))
        if "gpu" in it.keywords:
            if have_gpu is None:
                have_gpu = _cuda_ok()
            if not have_gpu:
                it.add_marker(pytest.mark.skip(reason="no CUDA device"))


@pytest.fixture(scope="session")
def ora():
    from oracle import refbind
    refbind.build(ref=os.path.isdir(refbind.REF_SRC))
    return refbind.oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("libref.so missing")
    return refbind.ref()
