"""TEST INFRASTRUCTURE ONLY -- golden vectors for the object layer (Quad / Series / Varf), produced by the
reference's UNMODIFIED Python package (/root/reference/hermipy) running on its own C++ (oracle/_ref, through
the hermite_cpp stand-in oracle/pyref/hermite_cpp.py).  matplotlib is not installed in this image and is
stubbed (the reference imports it at module top, series.py:27, varf.py:31; no plotting is done).

    python oracle/make_golden_api.py        ->  tests/golden/api.npz   (needs /root/reference; run in the build container)

The cases themselves live in tests/api_cases.py and are the same code the parity tests run against
hermipy_b200.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref"))
sys.path.insert(0, "/root/reference")

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.colors"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].rc = lambda *a, **k: None
sys.modules["matplotlib.ticker"].MaxNLocator = object
sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]

import numpy as np  # noqa: E402
import hermipy as hm  # noqa: E402   (the reference)
import api_cases  # noqa: E402

assert hm.__file__.startswith("/root/reference/"), hm.__file__
hm.settings['cachedir'] = '/tmp/hermipy_golden_cache'
out = api_cases.run(hm)
path = os.path.join(ROOT, "tests", "golden", "api.npz")
np.savez_compressed(path, **out)
print("wrote %s: %d arrays, %.1f kB" % (path, len(out), os.path.getsize(path) / 1e3))
