import ctypes
import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    try:
        cu = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    if cu.cuInit(0) != 0:
        return 0
    n = ctypes.c_int(0)
    if cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
        return 0
    return n.value


HAVE_GPU = _cuda_device_count() > 0


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def hr():
    """The product package (ctypes over the C ABI)."""
    import hr_pkg
    return hr_pkg.load()


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (checker)."""
    from oracle import hr_oracle_py
    hr_oracle_py.build()
    return hr_oracle_py


@pytest.fixture(scope="session")
def synth(hr):
    from hopperrender_b200 import synth as s
    return s
