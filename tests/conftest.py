"""pytest config: `-m gpu` tests need a B200 and call the CUDA path through the C-ABI;
`-m "not gpu"` tests cover the oracle, the goldens, the host logic and the ABI surface."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build the C-ABI library (nvcc cross-compiles without a GPU) and the C oracle once."""
    from cvcs_b200 import build as b
    b.build()
    import oracle
    oracle.build()
