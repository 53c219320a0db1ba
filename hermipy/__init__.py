"""Alias package: `import hermipy as hm` resolves to hermipy_b200, so scripts written against the
reference (urbainvaes/hermipy) run unchanged on the B200 backend.  Sub-modules are aliased one to one
(`hermipy.core`, `hermipy.quad`, ...)."""
import importlib
import sys

import hermipy_b200 as _impl
from hermipy_b200 import *  # noqa: F401,F403
from hermipy_b200 import Function, Operator, Position, Quad, Series, Varf, cache, core, settings, x, y, z, f  # noqa: F401

for _name in ("core", "quad", "series", "varf", "function", "operator", "position", "cache", "stats", "settings",
              "lib"):
    _module = importlib.import_module("hermipy_b200." + _name)
    sys.modules[__name__ + "." + _name] = _module
    if _name not in ("cache", "settings"):      # hm.cache / hm.settings are the decorator / the dictionary
        globals()[_name] = _module
del _name, _module

__version__ = _impl.__version__
