"""CPU tests (-m "not gpu") of the drop-in boundary: the C-ABI library loads, exports every symbol
include/hermite_b200.h declares, and the host-only index API is bit-exact against the golden vectors
generated from the compiled reference.  No compute entry point is exercised without a GPU -- except to
check that it fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SETS = ["triangle", "cube", "cross", "cross_nc", "rectangle"]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hermite_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hermipy_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build() / make -C hermipy_b200/csrc"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), name
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_index_api_matches_golden(core):
    g = np.load(os.path.join(GOLD, "index_sets.npz"))
    for key in g.files:
        if key.startswith("sizes"):
            continue
        s, d, N = key.rsplit("_", 2)
        d, N = int(d[1:]), int(N[1:])
        got = core.iterator_list_indices(d, N, s)
        assert np.array_equal(got, g[key]), key
        assert core.iterator_size(d, N, s) == len(g[key])
        assert core.iterator_get_degree(d, len(g[key]), s) == (N if not (s == "cross_nc" and N == 0) else 0)
    for key, val in zip(g["sizes_keys"], g["sizes_vals"]):
        s, d, N = str(key).rsplit("_", 2)
        assert core.iterator_size(int(d[1:]), int(N[1:]), s) == int(val), key


def test_closed_form_indices(core):
    for d in (1, 2, 3, 4):
        for s, N in (("triangle", 6), ("cube", 4)):
            for i, m in enumerate(core.iterator_list_indices(d, N, s)):
                assert core.iterator_index(m, s) == i
    # survey sizes
    assert core.iterator_size(2, 200) == 20301 and core.iterator_size(2, 400) == 80601
    assert core.iterator_size(3, 255) == 2829056 and core.iterator_size(2, 400, "cube") == 160801
    assert core.iterator_size(2, 150, "rectangle") == 11476


def test_errors_are_exceptions_not_exit(core):
    from hermipy_b200 import DegreeError, HermiteB200Error
    with pytest.raises(ValueError):
        core.iterator_size(2, 3, "sphere")
    with pytest.raises(DegreeError):          # reference: "Can't find x" + exit(0), lib.cpp:93-99
        core.iterator_get_degree(2, 7, "triangle")
    with pytest.raises(HermiteB200Error):     # reference: assert(degree > 0), iterators.cpp:328
        core.iterator_size(2, 0, "cross")
    # degrees above the reference's DEGREE_BISSECT_MAX (iterators.cpp:30) are fine here
    assert core.iterator_get_degree(2, core.iterator_size(2, 1000), "triangle") == 1000


def test_no_cpu_fallback(core):
    """Without a CUDA device the compute path must fail loudly (never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    from hermipy_b200 import HermiteB200Error
    x = np.linspace(-1, 1, 5)
    with pytest.raises(HermiteB200Error, match="CUDA"):
        core.transform(2, x, [x], [np.ones(5) / 5], True)
    with pytest.raises(HermiteB200Error, match="CUDA"):
        core.triple_products(3)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hermipy_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("oracle/", "oracle/").lower() or fn == "README.md", \
                    "%s mentions the oracle" % os.path.join(dirpath, fn)


def test_gauss_hermite_generator_beyond_numpy_limit():
    """hermipy/lib.py:24-33 replacement: correct for n >= 371 where numpy's hermegauss breaks (SURVEY finding 3)."""
    import scipy.special as sp
    from hermipy_b200.lib import hermegauss, hermegauss_nd
    for n in (1, 2, 7, 128, 401, 1001):
        x, w = hermegauss(n)
        assert len(x) == n and np.all(np.diff(x) > 0) and np.all(w >= 0) and np.all(np.isfinite(w))
        assert abs(w.sum() - 1) < 1e-14
        if n >= 3:
            assert abs(np.sum(w * x ** 2) - 1) < 1e-13 and abs(np.sum(w * x ** 4) - 3) < 1e-12   # tests/test_quad.py:39-51
        if n > 1:
            xs, ws = sp.roots_hermitenorm(n)
            ws = ws / np.sqrt(2 * np.pi)
            assert np.abs(x - xs).max() < 1e-12
            big = ws > 1e-280
            assert np.max(np.abs(w[big] - ws[big]) / ws[big]) < 1e-11
    nodes, weights = hermegauss_nd([3, 5])
    assert [len(n) for n in nodes] == [3, 5] and [len(w) for w in weights] == [3, 5]
