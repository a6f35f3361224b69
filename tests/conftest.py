import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# compiled patch kernels of the test runs stay inside the repository (the library's default is ~/.cache/maxib200); the directory does not
# travel to the GPU box (.gpurunignore), so the run-time compiler is exercised there
os.environ.setdefault("MXB_PATCH_CACHE", os.path.join(ROOT, ".mxb_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement oracle (always buildable: gcc only)."""
    from oracle import oracle_py
    oracle_py.build("port")
    return oracle_py


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled from /root/reference (absent on the GPU box unless prebuilt)."""
    from oracle import oracle_py
    if os.path.isdir("/root/reference/src"):
        oracle_py.build("reference")
    if not oracle_py.available("reference"):
        pytest.skip("oracle/_ref/libmaxiref.so not built and /root/reference absent")
    return oracle_py
