import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    gdir = ROOT / "tests" / "golden"

    def load(name):
        return np.load(gdir / name, allow_pickle=False)

    return load


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from groopm_b200 import _lib

    _lib.lib()  # the in-tree libgmdist.so must load: fail loudly otherwise
    return torch.device("cuda", 0)
