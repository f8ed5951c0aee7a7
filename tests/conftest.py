import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    data = np.load(os.path.join(GOLDEN_DIR, "golden_arrays.npz"))
    return {k.replace("__", "/"): data[k] for k in data.files}


@pytest.fixture(scope="session")
def golden_digests():
    out = {}
    with open(os.path.join(GOLDEN_DIR, "golden_digests.txt")) as f:
        for line in f:
            k, v = line.rstrip("\n").rsplit(" ", 1)
            out[k] = v
    return out


@pytest.fixture(scope="session")
def golden_state_keys():
    out = {}
    with open(os.path.join(GOLDEN_DIR, "golden_state_keys.txt")) as f:
        for line in f:
            case, key, shape = line.split()
            out.setdefault(case, {})[key] = shape
    return out
