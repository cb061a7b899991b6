import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) the in-tree shared library; CPU-only: nvcc cross-compiles sm_100a."""
    from emrt_b200 import build as b
    return b.build()


@pytest.fixture(scope="session")
def cuda_dev(lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from emrt_b200 import ops
    ops.device_check()          # fails loudly on anything but sm_100
    return torch.device("cuda:0")
