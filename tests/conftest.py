import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def lt_lib():
    """The C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    from locotouch_b200.csrc import build

    build.build()
    from locotouch_b200 import _C

    return _C.lib()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
