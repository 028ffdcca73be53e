import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle if they are missing (both build without a GPU)."""
    lib = os.path.join(ROOT, "glome_b200", "_build", "libglomecuda.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "glome_b200", "csrc")])
    orc = os.path.join(ROOT, "oracle", "_build", "libglome_oracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield


def has_gpu():
    try:
        import glome_b200._lib as L
        return L.load().glome_device_count() > 0
    except Exception:
        return False
