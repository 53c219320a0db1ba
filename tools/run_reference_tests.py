"""Developer check (needs /root/reference, CPU only): run the reference's OWN unittest files
(/root/reference/tests/test_{quad,series,tensorize_project,function,equations}.py, unmodified, read in place)
against this repository's object layer (`hermipy` alias -> hermipy_b200) with the compute functions of
hermipy_b200.core swapped for the reference's own C++ (tests/oracle_backend.py).  This validates the host-side
API (Quad/Series/Varf/Function/Operator/tensorize_at/cache) independently of the GPU.  `hermipy.equations` is an
application module outside this repository's scope: it is loaded from the reference tree for the run.

    python tools/run_reference_tests.py [test_quad test_series ...]
"""
import importlib.util
import os
import sys
import types
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"

import hermipy  # noqa: E402  (alias of hermipy_b200)
import oracle_backend  # noqa: E402

# matplotlib is not installed here: the reference's test files import it at module top
if importlib.util.find_spec("matplotlib") is None:
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.rc = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")

spec = importlib.util.spec_from_file_location("hermipy.equations", os.path.join(REF, "hermipy", "equations.py"))
eq = importlib.util.module_from_spec(spec)
sys.modules["hermipy.equations"] = eq
spec.loader.exec_module(eq)
hermipy.equations = eq

names = sys.argv[1:] or ["test_function", "test_series", "test_tensorize_project", "test_quad", "test_equations"]
suite = unittest.TestSuite()
for name in names:
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(mod))
with oracle_backend.oracle_core():
    result = unittest.TextTestRunner(verbosity=1).run(suite)
sys.exit(0 if result.wasSuccessful() else 1)
