import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def emu_lib():
    """g++ -DCRANE_EMU build of the kernel source (CPU kernel-emulation harness)."""
    from cranesched_b200.build import build_emu
    return build_emu()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; never built lazily on the GPU box's behalf of a fallback."""
    from cranesched_b200.scheduler import LIB_PATH
    if not os.path.exists(LIB_PATH):
        from cranesched_b200.build import build_cuda
        build_cuda()
    return LIB_PATH
