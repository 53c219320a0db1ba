"""hermipy_b200 -- B200-native (sm_100a CUDA) implementation of hermipy's numerical core behind
hermipy's own Python API: `Quad`, `Series`, `Varf`, `Function`, `Operator`, `Position`, `cache`,
`settings` (reference: hermipy/__init__.py:21-36), with every numerical call going through the C ABI of
libhermite_b200.so (hermipy_b200.core).  `import hermipy_b200 as hm` is the counterpart of
`import hermipy as hm`; the repository root also carries a `hermipy` alias package so that existing
scripts run unchanged.  See DESIGN.md and INTEGRATION.md."""
from . import core  # noqa: F401
from ._lib import HermiteB200Error, DegreeError, LIB_PATH  # noqa: F401

# classes in the package namespace
from .function import Function  # noqa: F401
from .operator import Operator  # noqa: F401
from .position import Position  # noqa: F401
from .series import Series  # noqa: F401
from .varf import Varf  # noqa: F401
from .quad import Quad  # noqa: F401

# the cache decorator and the settings dictionary
from .cache import cache  # noqa: F401
from .settings import settings  # noqa: F401

# for convenience
x, y, z = Function.xyz[0:3]
f = Operator.f

__version__ = "0.2.0"
