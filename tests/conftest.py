import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_usable():
    try:
        import torch
        if not torch.cuda.is_available():
            return False
        import cordic_b200
        return cordic_b200.lib().zc_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a GPU-less machine skips the gpu tier instead of failing in torch.cuda init."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items or _gpu_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (the gpu tier runs on the B200 box)")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle (checker) and the product library exist; both build without a GPU."""
    oracle_so = os.path.join(ROOT, "oracle", "libzc_oracle.so")
    if not os.path.exists(oracle_so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    import cordic_b200
    cordic_b200.lib()
    yield


def has_reference():
    return os.path.exists("/root/reference/rtl/cordic.v")
