import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu under gpurun")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gold():
    import numpy as np
    return lambda name: np.load(os.path.join(GOLD, name))


@pytest.fixture(scope="session")
def checksum():
    def f(sd):
        return float(sum(float(v.double().abs().sum()) * (1 + (i % 7)) for i, (k, v) in enumerate(sorted(sd.items()))))
    return f
