import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from realtimeparticles_b200 import _abi
        return _abi.lib().rtp_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """tests never run against stale binaries: (re)build the oracle and the CUDA library if sources are newer."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    lib = os.path.join(ROOT, "realtimeparticles_b200", "lib", "librtp_cuda.so")
    if not os.path.exists(lib) or os.path.exists("/usr/local/cuda/bin/nvcc"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "realtimeparticles_b200", "csrc"), "-s", "-j8"], check=True)
    yield
