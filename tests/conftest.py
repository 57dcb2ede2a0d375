import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unhex(v):
    if isinstance(v, str):
        try:
            return float.fromhex(v)
        except ValueError:
            return v
    if isinstance(v, list):
        return [unhex(u) for u in v]
    if isinstance(v, dict):
        return {k: unhex(u) for k, u in v.items()}
    return v


@pytest.fixture(scope="session")
def gold():
    with open(os.path.join(GOLDEN, "reference_outputs.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def gold_windows():
    with open(os.path.join(GOLDEN, "windows.json")) as fh:
        return json.load(fh)


def rel_close(a, b, tol=1e-12):
    """Relative comparison used for every fp64 statistic (north star: 1e-12)."""
    if a != a or b != b:
        return a != a and b != b
    if a == b:
        return True
    return abs(a - b) <= tol * max(abs(a), abs(b))
