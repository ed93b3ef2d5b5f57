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
    """CPU oracle (test infrastructure): builds oracle/libhp3d_oracle.so on first use."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def gpulib():
    """The product's C-ABI library, loaded through ctypes (host-only entry points work without a GPU)."""
    from hp3d_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def gpu(gpulib):
    from hp3d_b200 import _lib
    _lib.check(gpulib.hp3d_gpu_init(0))
    return gpulib
