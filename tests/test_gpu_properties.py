"""GPU tests in the style of the reference's own suite (identities and analytic answers through the
public function set, tests/test_quad.py and tests/test_tensorize_project.py of the reference), plus
parity at the BASELINE.json config shapes on bounded samples, edge cases and error behaviour."""
import numpy as np
import numpy.polynomial.hermite_e as herm
import pytest
from math import factorial

from conftest import gauss_hermite, rel_err

pytestmark = pytest.mark.gpu


# ---- reference tests/test_quad.py:37-108 (TestIntegrate) ------------------------------------------
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_integrate_moments(core, dim):
    x, w = gauss_hermite(64)
    nodes, weights = [x] * dim, [w] * dim
    grids = np.meshgrid(*nodes, indexing="ij")
    assert abs(core.integrate(np.ones_like(grids[0]), nodes, weights) - 1) < 1e-12      # :39-45
    assert abs(core.integrate(grids[0], nodes, weights)) < 1e-12                        # mean
    assert abs(core.integrate(grids[-1] ** 2, nodes, weights) - 1) < 1e-11              # variance
    assert abs(core.integrate(grids[0] ** 4, nodes, weights) - 3) < 1e-10               # 4th moment


# ---- tests/test_quad.py:111-162 (TestHermiteTransform) ---------------------------------------------
def test_transform_of_one_is_first_unit_vector_3d(core):
    x, w = gauss_hermite(35)
    c = core.transform(30, np.ones((35, 35, 35)), [x] * 3, [w] * 3, True)              # :113-119
    e0 = np.zeros_like(c)
    e0[0] = 1
    assert np.abs(c - e0).max() < 1e-12


def test_eval_matches_numpy_hermeval(core):
    x, w = gauss_hermite(40)                                                           # :121-133
    degree = 25
    for k in (0, 1, 7, degree):
        c = np.zeros(degree + 1)
        c[k] = 1
        got = core.transform(degree, c, [x], [w], False)
        want = herm.hermeval(x, c) / np.sqrt(float(factorial(k)))
        assert rel_err(got, want) < 1e-12


@pytest.mark.parametrize("index_set", ["triangle", "cube", "cross", "rectangle"])
def test_forward_backward_identity_on_polynomials(core, index_set):
    x, w = gauss_hermite(60)                                                           # :135-142
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = 1 + X - 2 * Y + X * Y + 0.5 * X ** 2 * Y - Y ** 3 / 3
    degree = 10
    c = core.transform(degree, f, [x, x], [w, w], True, index_set=index_set)
    g = core.transform(degree, c, [x, x], [w, w], False, index_set=index_set).reshape(60, 60)
    if index_set in ("triangle", "cube"):          # the sparse sets do not contain x^2 y and y^3
        sw = np.sqrt(np.outer(w, w))
        assert rel_err(g * sw, f * sw) < 1e-12
    c2 = core.transform(degree, g, [x, x], [w, w], True, index_set=index_set)
    assert rel_err(c2, c) < 1e-12


# ---- tests/test_quad.py:165-192, 279-330 (varf) ------------------------------------------------------
@pytest.mark.parametrize("index_set", ["triangle", "cube"])
def test_varf_of_one_is_identity(core, index_set):
    x, w = gauss_hermite(61)
    A = core.varf(30, np.ones((61, 61)), [x, x], [w, w], index_set=index_set)          # :167-172
    assert np.abs(A - np.eye(len(A))).max() < 1e-12
    S = core.varf(30, np.ones((61, 61)), [x, x], [w, w], index_set=index_set, sparse=True)
    assert S.nnz == S.shape[0]                                                         # :292-296


def test_varf_bandwidth_equals_polynomial_degree(core):
    x, w = gauss_hermite(101)                                                          # :298-313
    for k in range(1, 6):
        S = core.varf(50, x ** k, [x], [w], sparse=True)
        rows, cols = S.nonzero()
        assert np.abs(rows - cols).max() == k
        assert rel_err(S.toarray(), core.varf(50, x ** k, [x], [w])) == 0.0            # sparse == dense (:315-320)


def test_ornstein_uhlenbeck_generator_is_diagonal(core):
    """<(-x d/dx + d^2/dx^2) phi_m, phi_n> = -m delta_mn: assembled from varf + varfd like Quad.discretize_op
    (quad.py:344-390), tests/test_quad.py:174-181."""
    degree = 40
    x, w = gauss_hermite(2 * degree + 1)
    one = core.varf(degree, np.ones_like(x), [x], [w])
    vx = core.varf(degree, x, [x], [w])
    d1 = core.varfd(1, degree, 0, vx)            # x d/dx
    d2 = core.varfd(1, degree, 0, core.varfd(1, degree, 0, one))
    L = -d1 + d2
    assert np.abs(L - np.diag(-np.arange(degree + 1.0))).max() < 1e-10


# ---- parity at the BASELINE.json shapes, bounded samples -------------------------------------------
def test_c1_degree100_transform_and_varf_vs_reference(core, ref):
    x, w = gauss_hermite(201)
    rng = np.random.default_rng(1)
    f = np.exp(-x ** 2 / 8) * np.cos(3 * x)
    assert rel_err(core.transform(100, f, [x], [w], True), ref.transform(100, f, [x], [w], True)) < 1e-12
    C = rng.standard_normal((8, 101)) / (1 + np.arange(101)) ** 2
    F = core.transform_batch(100, C, [x], [w], False)
    for b in range(8):
        assert rel_err(F[b] * np.sqrt(w), ref.transform(100, C[b], [x], [w], False) * np.sqrt(w)) < 1e-12
    for f in (x, x ** 2, x ** 4 / 4 - x ** 2 / 2, np.ones_like(x)):
        A, R = core.varf(100, f, [x], [w]), ref.varf(100, f, [x], [w])
        assert rel_err(A, R) < 1e-12 and np.array_equal(A != 0, R != 0)


def test_c5_degree1000_rows_vs_reference(core, ref):
    """Batched 1-D, degree 1000, on the nodes where the reference formula is finite."""
    x, w = gauss_hermite(2001)
    keep = (w > 0) & (np.abs(x) <= 50)
    x, w = x[keep], w[keep]
    rng = np.random.default_rng(5)
    C = rng.standard_normal((6, 1001)) / (1 + np.arange(1001))
    F = core.transform_batch(1000, C, [x], [w], False)
    assert np.all(np.isfinite(F))
    got = core.transform_batch(1000, F, [x], [w], True)
    for b in range(6):
        f_ref = ref.transform(1000, C[b], [x], [w], False)
        assert rel_err(F[b] * np.sqrt(w), f_ref * np.sqrt(w)) < 1e-12
        assert rel_err(got[b], ref.transform(1000, F[b], [x], [w], True)) < 1e-12


def test_c4_reduced_3d_cube_vs_reference(core, ref):
    n, degree = 20, 19
    x, w = gauss_hermite(n)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    f = sum(np.exp(-((X - a) ** 2 + (Y - b) ** 2 + (Z - c) ** 2) / 6) for a, b, c in [(0, 0, 0), (1, -1, .5), (-2, .3, 1)])
    c = core.transform(degree, f, [x] * 3, [w] * 3, True, index_set="cube")
    assert rel_err(c, ref.transform(degree, f.ravel(), [x] * 3, [w] * 3, True, index_set="cube")) < 1e-12


def test_c2_reduced_2d_degree60_vs_reference(core, ref):
    x, w = gauss_hermite(121)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = np.exp(-(X ** 2 + X * Y + Y ** 2) / 10)
    for s in ("cube", "triangle"):
        assert rel_err(core.transform(60, f, [x, x], [w, w], True, index_set=s),
                       ref.transform(60, f.ravel(), [x, x], [w, w], True, index_set=s)) < 1e-12
    V = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
    A = core.varf(30, V, [x, x], [w, w], index_set="triangle")
    R = ref.varf(30, V.ravel(), [x, x], [w, w], index_set="triangle")
    assert rel_err(A, R) < 1e-12 and np.array_equal(A != 0, R != 0)
    assert rel_err(core.varf(30, V, [x, x], [w, w], index_set="triangle", mode="gram"), R) < 1e-12


def test_degree_above_reference_cap(core, ref):
    """The reference's DEGREE_BISSECT_MAX 250 (iterators.cpp:30) is gone; checked against the oracle build
    with the cap raised (oracle/Makefile, libhermite_ref_big.so)."""
    if not ref.available("_big"):
        pytest.skip("libhermite_ref_big.so not built")
    rng = np.random.default_rng(9)
    v0, v1 = rng.standard_normal(301), rng.standard_normal(301)
    got = core.tensorize([v0, v1])
    want = ref.tensorize([v0, v1], variant="_big")
    assert np.array_equal(got, want)
    assert np.array_equal(core.project(got, 2, [1]), ref.project(want, 2, [1], variant="_big"))


# ---- edge cases and errors -----------------------------------------------------------------------------
def test_edge_cases(core, ref):
    x, w = gauss_hermite(7)
    one_node = (np.array([0.0]), np.array([1.0]))
    assert np.allclose(core.transform(0, [2.5], [one_node[0]], [one_node[1]], True), [2.5])          # n = 1, degree 0
    assert rel_err(core.transform(6, x ** 2, [x], [w], True), ref.transform(6, x ** 2, [x], [w], True)) < 1e-12
    # odd sizes everywhere (pitch padding), degree + 1 > number of nodes (rank deficient but well defined)
    y, v = gauss_hermite(5)
    f = np.outer(np.sin(x), np.cos(y))
    assert rel_err(core.transform(8, f, [x, y], [w, v], True), ref.transform(8, f.ravel(), [x, y], [w, v], True)) < 1e-12
    assert core.transform_batch(3, np.zeros((0, 7)), [x], [w], True).shape == (0, 4)                  # empty batch
    # zero weights contribute zero, not NaN
    w0 = w.copy()
    w0[0] = 0.0
    assert np.all(np.isfinite(core.transform(6, np.ones(7), [x], [w0], True)))
    # NaN input propagates like in the reference (not skipped by the varf threshold, varf.cpp:238-241)
    fn = x.copy()
    fn[3] = np.nan
    assert np.isnan(core.varf(3, fn, [x], [w])).all()


def test_errors(core):
    from hermipy_b200 import HermiteB200Error
    x, w = gauss_hermite(8)
    with pytest.raises(ValueError):
        core.transform(3, np.ones(9), [x], [w], True)                    # wrong input length
    with pytest.raises(HermiteB200Error):                               # transform.cpp:92-96: odd degree on a Fourier axis
        core.transform(3, np.ones(8), [x], [w], True, do_fourier=[1])
    with pytest.raises(HermiteB200Error):                               # tensorize.cpp:44-52: direction given twice
        core.tensorize({frozenset({0}): np.ones(4), frozenset({0, 1}): np.ones(10)})
    with pytest.raises(HermiteB200Error):                               # mismatching factor sizes (tensorize.cpp:72-81)
        core.tensorize([np.ones(4), np.ones(5)])
