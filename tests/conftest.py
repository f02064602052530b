import os
import sys

import numpy as np
import pytest

TESTS = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(TESTS)
PKG = os.path.join(REPO, "pytorch-quantity_b200")
GOLDEN = os.path.join(TESTS, "golden")
for p in (PKG, REPO, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


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


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_json(npz, key="json"):
    import json
    return json.loads(bytes(npz[key]).decode())


@pytest.fixture(scope="session")
def oracle():
    from oracle import pq_oracle
    pq_oracle.build()
    return pq_oracle
