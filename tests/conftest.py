import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "stereo-semantic-vo_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _cuda_devices():
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are skipped (not failed), so a plain `pytest tests` on a CPU box shows the
    state of the CPU suite; on a GPU box nothing is skipped and a missing library still fails loudly."""
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: libsvo_b200 has no CPU fallback (run on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
