import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load the C-ABI library; fails loudly when nvcc or the library is missing."""
    from selavi_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is available")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
