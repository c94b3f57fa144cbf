import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ic_golden():
    """Outputs of the unmodified reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, "ic_reference.npz"))
    names = sorted({k.split("__")[0] for k in z.files})
    return {n: (z[n + "__X"], z[n + "__C"], z[n + "__Y"]) for n in names}


def random_target(rng, K):
    """Recipe of reference tests/test_iman_conover.py:154-155."""
    A = rng.normal(size=(2 * K, K))
    return 0.9 * np.corrcoef(A, rowvar=False) + 0.1 * np.eye(K)
