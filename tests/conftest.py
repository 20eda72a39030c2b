import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """The C-ABI library (built in-tree if it is missing or stale)."""
    from yourmt3_b200 import build as B
    B.build_native()
    from yourmt3_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def emu_lib():
    import ctypes
    from yourmt3_b200 import build as B
    return ctypes.CDLL(B.build_host_emu())


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
