import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def cases_of(npz):
    return sorted({k.split("/")[0] for k in npz.files if "/" in k})


def rel_err(a, b):
    """max |a-b| / max(|b|) — the north-star parity metric (per tensor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.max(np.abs(b))), 1e-30) if b.size else 1.0
    return float(np.max(np.abs(a - b))) / denom if a.size else 0.0


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
