import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-locality-sensitive-hashing_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "nlsh_golden.npz")
    return dict(np.load(path, allow_pickle=False))


@pytest.fixture(scope="session")
def oracle():
    from oracle import nlsh_oracle
    return nlsh_oracle
