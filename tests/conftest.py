import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def gauss_hermite(n):
    """Probabilists' Gauss-Hermite rule normalised to total weight 1, as hermipy/lib.py:24-33 builds it
    (scipy instead of numpy.hermegauss, which returns zero/NaN weights for n >= 371 with NumPy 2.3)."""
    import scipy.special as sp
    x, w = sp.roots_hermitenorm(n)
    return x, w / np.sqrt(2 * np.pi)


def rel_err(a, b):
    """norm-wise relative error ||a-b||_2 / max(||b||_2, tiny)"""
    a, b = np.asarray(a, dtype=float).ravel(), np.asarray(b, dtype=float).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as _ref
    if not _ref.available():
        pytest.skip("oracle/_ref/libhermite_ref.so not built (run `make -C oracle ref` where /root/reference exists)")
    return _ref


@pytest.fixture(scope="session")
def core():
    from hermipy_b200 import core as _core
    return _core
