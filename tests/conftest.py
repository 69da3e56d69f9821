import importlib
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


def load_pkg():
    """The package directory name starts with a digit, so it is imported through importlib."""
    return importlib.import_module("3deecelltracker_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def split_cases(npz):
    cases = {}
    for key in npz.files:
        c, k = key.split("__", 1)
        cases.setdefault(c, {})[k] = npz[key]
    return cases
