import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def vlb():
    m = importlib.import_module("vulkan-light-bakery_b200")
    if not os.path.exists(m.LIB_PATH):
        importlib.import_module("vulkan-light-bakery_b200.build").build()
    return m


@pytest.fixture(scope="session")
def scenes(vlb):
    return importlib.import_module("vulkan-light-bakery_b200.scenes")


@pytest.fixture(scope="session")
def oa():
    from oracle import build_oracle, oracle_api
    build_oracle.build_oracle()
    build_oracle.build_ref()
    return oracle_api


@pytest.fixture(scope="session")
def ctx(vlb):
    c = vlb.Context(0)
    yield c
    c.close()


def rel_l2(a, b):
    """max over SH vectors of ||a-b|| / ||b|| (the tolerance metric of BASELINE.json)."""
    import numpy as np
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    a = a.reshape(-1, 48)
    b = b.reshape(-1, 48)
    den = np.linalg.norm(b, axis=1)
    den = np.where(den > 0, den, 1.0)
    return float((np.linalg.norm(a - b, axis=1) / den).max())
