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


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session (no-op when up to date)."""
    from sdfibm_b200 import build

    build.build()
    build.build_mesh()
    build.build_host()
    build.build_oracle()


@pytest.fixture(scope="session")
def m1_points():
    return np.load(os.path.join(GOLDEN, "m1_points.npz"))["points"]


@pytest.fixture(scope="session")
def m2_points():
    return np.load(os.path.join(GOLDEN, "m2_points.npz"))["points"]


@pytest.fixture(scope="session")
def g1_alpha():
    return np.load(os.path.join(GOLDEN, "g1_alpha.npz"))["alpha"]


@pytest.fixture(scope="session")
def g2_As():
    return np.load(os.path.join(GOLDEN, "g2_As.npz"))["As_central"]
