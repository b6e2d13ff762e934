import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "esn-ofdm-mimo_b200")
for p in (ROOT, os.path.join(ROOT, "tests", "golden"), PKG, os.path.join(PKG, "libs")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64) if not np.iscomplexobj(a) else np.asarray(a)
    b = np.asarray(b, dtype=np.float64) if not np.iscomplexobj(b) else np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / (np.linalg.norm(b.ravel()) + 1e-300))
