"""GPU tests (-m gpu) of the fused object-API entry points: Quad.transform / transform_batch / eval / varf in one device
pass each (hermipy/quad.py:287-342 does the discretisation and the weighting with NumPy on the host).  Compared bit for
bit with the same steps composed on the host from the separate C-ABI calls, and with the reference."""
import numpy as np
import pytest

from conftest import gauss_hermite, rel_err

pytestmark = pytest.mark.gpu


def _setup(n=36):
    x, w = gauss_hermite(n)
    mean, dil = np.array([0.3, -0.2]), np.array([[1.1, 0.0], [0.0, 0.7]])
    f_body = "exp(-v[0]*v[0]/5 - v[1]*v[1]/3) * (1 + v[0]*v[1])"
    fac_body = "exp(-v[0]*v[0]/4 - v[1]*v[1]/4)"
    return x, w, mean, dil, f_body, fac_body


@pytest.mark.parametrize("index_set", ["triangle", "cube"])
def test_transform_weighted_equals_the_host_composition(core, ref, index_set):
    x, w, mean, dil, f_body, fac_body = _setup()
    nodes, weights = [x, x], [w, w]
    f = core.discretize(f_body, nodes, mean, dil)
    fac = core.discretize(fac_body, nodes, mean, dil)
    mapped = f / (1e-300 * (fac == 0) + fac)
    want = core.transform(12, mapped, nodes, weights, True, index_set=index_set)
    # expression in, coefficients out; grid values in; a batch of grid values in
    got_expr = core.transform_weighted(12, f_body, fac_body, nodes, weights, mean, dil, index_set=index_set)
    got_grid = core.transform_weighted(12, f, fac_body, nodes, weights, mean, dil, index_set=index_set)
    assert np.array_equal(got_expr, want) and np.array_equal(got_grid, want)
    batch = np.stack([f, 2 * f, f * f])
    got_b = core.transform_weighted(12, batch, fac_body, nodes, weights, mean, dil, index_set=index_set)
    assert got_b.shape == (3, len(want)) and np.array_equal(got_b[0], want)
    assert np.array_equal(got_b[2], core.transform(12, f * f / (1e-300 * (fac == 0) + fac), nodes, weights, True, index_set=index_set))
    # no factor; and the reference's own transform of the same mapped values
    assert np.array_equal(core.transform_weighted(12, f, None, nodes, weights, mean, dil, index_set=index_set),
                          core.transform(12, f, nodes, weights, True, index_set=index_set))
    assert rel_err(got_expr, ref.transform(12, ref.discretize(f_body, nodes, mean, dil) / ref.discretize(fac_body, nodes, mean, dil),
                                           nodes, weights, True, index_set=index_set)) <= 1e-12


def test_factor_that_underflows_divides_by_1e_300(core):
    x, w = gauss_hermite(60)
    fac_body = "exp(-40*v[0]*v[0])"            # 0 at the outer nodes in double precision
    f = np.exp(-45 * x * x)
    fac = core.discretize(fac_body, [x], [0.0], [[1.0]])
    assert (fac == 0).any()
    want = core.transform(8, f / (1e-300 * (fac == 0) + fac), [x], [w], True)
    assert np.array_equal(core.transform_weighted(8, f, fac_body, [x], [w], [0.0], [[1.0]]), want)


def test_eval_weighted_equals_the_host_composition(core):
    x, w, mean, dil, f_body, fac_body = _setup()
    nodes, weights = [x, x], [w, w]
    coeffs = core.transform_weighted(10, f_body, fac_body, nodes, weights, mean, dil)
    mapped_nodes = [x * 0.9 + 0.1, x * 1.2 - 0.05]
    values = core.transform(10, coeffs, mapped_nodes, weights, False)
    want = values * core.discretize(fac_body, nodes, mean, dil)
    got = core.eval_weighted(10, coeffs, fac_body, mapped_nodes, nodes, weights, mean, dil)
    assert np.array_equal(got, want)
    assert np.array_equal(core.eval_weighted(10, coeffs, None, mapped_nodes, nodes, weights, mean, dil), values)


@pytest.mark.parametrize("sparse", [False, True])
def test_varf_function_equals_varf_of_the_discretised_function(core, sparse):
    x, w, mean, dil, _, _ = _setup(30)
    nodes, weights = [x, x], [w, w]
    body = "v[0]*v[0]*v[0]*v[0]/4 - v[0]*v[0]/2 + v[1]*v[1]/2 + v[0]*v[1]/2"
    want = core.varf(9, core.discretize(body, nodes, mean, dil), nodes, weights, sparse=sparse, index_set="triangle")
    got = core.varf_function(9, body, nodes, weights, mean, dil, sparse=sparse, index_set="triangle")
    if sparse:
        assert (got != want).nnz == 0 and np.array_equal(got.indices, want.indices)
    else:
        assert np.array_equal(got, want)


def test_quad_methods_use_the_fused_calls():
    """Quad.transform of an expression makes no separate discretize / transform call (one fused C call)"""
    import hermipy_b200 as hm
    q = hm.Quad.gauss_hermite(20, dim=2, mean=[0.1, 0.2], cov=[[1.0, 0], [0, 2.0]])
    s = q.transform("exp(-x*x/3 - y*y/3)", degree=8, tensorize=False)   # (tensorize=True splits into 1-D transforms)
    vals = q.eval(s)
    A = q.varf("x*x + y", degree=6)
    assert np.isfinite(s.coeffs).all() and np.isfinite(vals).all() and A.matrix.shape[0] == 28
    x, w = q.nodes, q.weights
    f = q.discretize("exp(-x*x/3 - y*y/3)")
    fac = q.discretize(q.factor)
    want = hm.core.transform(8, f / (1e-300 * (fac == 0) + fac), x, w, True)
    assert np.array_equal(s.coeffs, want)
