import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(got, ref, floor_frac=1e-12):
    """max |got - ref| / max(|ref|, floor), floor = floor_frac * max|ref|."""
    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    scale = np.max(np.abs(ref)) if ref.size else 0.0
    floor = max(floor_frac * scale, 1e-300)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor))) if ref.size else 0.0


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture()
def ctx():
    from pybo_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()
