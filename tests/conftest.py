import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Load tests/golden/<name>.npz into a nested dict ('a/b' keys become d['a']['b'])."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        if "/" in k:
            a, b = k.split("/", 1)
            out.setdefault(a, {})[b] = z[k]
        else:
            out[k] = z[k]
    return out


@pytest.fixture(scope="session")
def flower_sd():
    return load_golden("flower_weights")["sd"]
