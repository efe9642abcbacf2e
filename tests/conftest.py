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
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("matching", "svd_decompose", "eight_point")}


def e_dist(Ea, Eb):
    """Frobenius distance after sign and scale normalisation (north_star's E metric)."""
    Ea = np.asarray(Ea, np.float64).reshape(-1)
    Eb = np.asarray(Eb, np.float64).reshape(-1)
    Ea = Ea / np.linalg.norm(Ea)
    Eb = Eb / np.linalg.norm(Eb)
    return min(np.linalg.norm(Ea - Eb), np.linalg.norm(Ea + Eb))
