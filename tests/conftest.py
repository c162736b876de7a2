import os
import sys

import pytest

os.environ.setdefault("T3D_DEBUG_CHECKS", "1")     # verify caller promises (thermal_replicated) on the device in tests
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    """The C-ABI library must exist (built by __graft_entry__.build())."""
    from thermal3d_vision_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib


@pytest.fixture(scope="session")
def cuda_device(lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    lib_built.lib()          # loads libt3d_sm100.so into this process; raises if missing
    return torch.device("cuda:0")
