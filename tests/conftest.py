import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "2023-compact-image-compression_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name is not a Python identifier, hence importlib)."""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "golden_small.npz")))


@pytest.fixture(scope="session")
def golden_digests():
    import json
    return json.load(open(os.path.join(GOLDEN, "golden_digests.json")))


@pytest.fixture(scope="session")
def golden_cct():
    return open(os.path.join(GOLDEN, "reference_testing.cct"), "rb").read()
