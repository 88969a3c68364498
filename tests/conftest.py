import json
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


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_batches():
    return np.load(os.path.join(GOLDEN, "batches.npz"))


@pytest.fixture(scope="session")
def golden_datasets():
    out = {}
    for name in ("rev", "fwd"):
        with open(os.path.join(GOLDEN, "dataset_%s.json" % name)) as f:
            out[name] = json.load(f)
    return out
