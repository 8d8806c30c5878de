import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device (this container, CI) the GPU tests are skipped, not failed.  A missing libb2s.so is NOT a reason
    to skip: on a GPU box that has to fail loudly."""
    try:
        from calibrating_b200 import _ffi
        ndev = _ffi.lib().b2s_device_count()
    except Exception:
        return
    if ndev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (b2s_device_count() == 0)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
