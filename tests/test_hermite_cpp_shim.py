"""The `hermite_cpp` module at the repository root (the name of the reference's Boost.Python bridge,
module.cpp:45-168) over libhermite_b200.so.  CPU: the converter surface of module.cpp:123-135 and, where the
reference is present, the reference's own tests/test_core.py run unchanged on top of it.  GPU: the reference-style
calls (Boost.Python argument order, container classes) against hermipy_b200.core."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as ss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def test_converters_round_trip():
    import hermite_cpp as hm_cpp
    a = np.random.default_rng(0).random((40, 30))
    for conv in (hm_cpp.to_mat, hm_cpp.to_bmat, hm_cpp.to_boost_mat):
        assert np.array_equal(hm_cpp.to_numpy(conv(a)), a)
    m = ss.random(12, 9, density=0.3, random_state=1, format="csr")
    sp = hm_cpp.to_spmat(list(m.data), list(m.indices), list(m.indptr), 12, 9)
    assert sp.size1() == 12 and sp.size2() == 9
    assert np.array_equal(hm_cpp.to_numpy(hm_cpp.full(sp)), m.toarray())
    r, c, v = hm_cpp.row_col_val(sp)
    assert np.array_equal(ss.csr_matrix((v, (r, c)), shape=(12, 9)).toarray(), m.toarray())
    vec = hm_cpp.double_vec()
    vec.extend([1.0, 2.0])
    assert np.array_equal(hm_cpp.to_numpy(vec), [1.0, 2.0])


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only present in the build container")
def test_reference_test_core_runs_unchanged_on_the_shim():
    """python -m unittest of the reference's tests/test_core.py, its own hermipy/core.py imported from the reference
    tree, `hermite_cpp` resolved to this repository's module (the converters are host code: no GPU needed)"""
    code = ("import sys, types; sys.path.insert(0, %r); import hermite_cpp; sys.path.insert(0, %r); "
            "[sys.modules.setdefault(n, types.ModuleType(n)) for n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'matplotlib.colors')]; "
            "mpl = sys.modules['matplotlib']; mpl.use = mpl.rc = lambda *a, **k: None; "
            "sys.modules['matplotlib.ticker'].MaxNLocator = object; mpl.ticker = sys.modules['matplotlib.ticker']; "
            "mpl.colors = sys.modules['matplotlib.colors']; mpl.pyplot = sys.modules['matplotlib.pyplot']; "
            "import unittest, importlib.util; "
            "spec = importlib.util.spec_from_file_location('ref_test_core', %r); m = importlib.util.module_from_spec(spec); "
            "spec.loader.exec_module(m); "
            "r = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(m)); "
            "sys.exit(0 if r.wasSuccessful() else 1)" % (ROOT, REFERENCE, os.path.join(REFERENCE, "tests", "test_core.py")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd="/tmp")
    assert out.returncode == 0, out.stderr[-3000:]


@pytest.mark.gpu
def test_reference_style_calls_on_the_gpu():
    import hermite_cpp as hm_cpp
    from conftest import gauss_hermite
    from hermipy_b200 import core
    x, w = gauss_hermite(24)
    nodes, weights = hm_cpp.double_mat(), hm_cpp.double_mat()
    for _ in range(2):
        nv, wv = hm_cpp.double_vec(), hm_cpp.double_vec()
        nv.extend(x)
        wv.extend(w)
        nodes.append(nv)
        weights.append(wv)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = (X * X * Y + 1).ravel()
    fv = hm_cpp.double_vec()
    fv.extend(f)
    four = hm_cpp.int_vec()
    four.extend([0, 0])
    c = hm_cpp.transform(8, fv, nodes, weights, four, True, "triangle")
    assert np.array_equal(np.array(c), core.transform(8, f, [x, x], [w, w], True))
    A = hm_cpp.varf(6, fv, nodes, weights, four, "cube")
    assert np.array_equal(hm_cpp.to_numpy(A), core.varf(6, f, [x, x], [w, w], index_set="cube"))
    S = hm_cpp.varf_sp(6, fv, nodes, weights, four, "cube")
    r, cc, v = hm_cpp.row_col_val(S)
    assert np.array_equal(ss.csr_matrix((v, (np.array(r, int), np.array(cc, int))), shape=A.a.shape).toarray(), A.a)
    D = hm_cpp.varfd(2, 6, 0, S, 0, "cube")
    assert np.array_equal(hm_cpp.to_numpy(hm_cpp.full(D)), core.varfd(2, 6, 0, A.a, index_set="cube"))
    T = hm_cpp.tensorize_sp([hm_cpp.to_boost_mat(np.eye(7)), hm_cpp.to_boost_mat(np.diag(np.arange(1.0, 8)))], "cube")
    assert T.m.nnz == 49 and hm_cpp.cube_size(2, 6) == 49
