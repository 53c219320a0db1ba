"""hermipy_b200 -- B200-native (sm_100a CUDA) implementation of the numerical core of hermipy:
Hermite transform, Hermite-function recurrence and varf / triple-product Galerkin assembly, behind
the same `core` function set as the reference's hermipy/core.py.  See DESIGN.md and INTEGRATION.md."""
from . import core  # noqa: F401
from ._lib import HermiteB200Error, DegreeError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
