"""Resident (in-HBM) varf matrices and their consumers: hb_dev_matvec / hb_dev_varfd through the C ABI and
`ResidentVarf` against the host `Varf` of the same operator (reference semantics: hermipy/varf.py:127-237)."""
import numpy as np
import pytest
import sympy as sym

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,ld,off", [(1, 1, 1, 0), (5, 7, 7, 0), (101, 101, 101, 0), (257, 1000, 1000, 0),
                                        (64, 999, 1002, 1), (9, 4097, 4099, 2), (2050, 2051, 2051, 0),
                                        (3, 2, 8, 3), (4001, 513, 514, 0)])
def test_matvec_matches_numpy(M, N, ld, off):
    import torch
    from hermipy_b200.resident import device_matvec
    rng = np.random.default_rng(M * 7919 + N)
    store = rng.standard_normal((M, ld + 4))
    A_host = store[:, off:off + N]
    x = rng.standard_normal(N)
    store_d = torch.from_numpy(store).cuda()
    A = store_d[:, off:off + N]                      # row stride ld + 4, base misaligned by `off` doubles
    guard = torch.full((M + 8,), 7.0, dtype=torch.float64, device="cuda")
    y = device_matvec(A, torch.from_numpy(x).cuda(), out=guard[4:4 + M])
    want = A_host @ x
    scale = np.abs(A_host) @ np.abs(x)
    assert np.all(np.abs(y.cpu().numpy() - want) <= 4e-16 * np.log2(N + 2) * scale + 1e-300)
    assert torch.all(guard[:4] == 7.0) and torch.all(guard[4 + M:] == 7.0)
    # reproducible run to run (fixed summation tree)
    y2 = device_matvec(A, torch.from_numpy(x).cuda())
    assert torch.equal(y2, guard[4:4 + M])


def test_matvec_propagates_nan_and_rejects_bad_shapes():
    import torch
    from hermipy_b200.resident import device_matvec
    A = torch.ones((6, 6), dtype=torch.float64, device="cuda")
    A[2, 5] = float("nan")
    y = device_matvec(A, torch.ones(6, dtype=torch.float64, device="cuda")).cpu().numpy()
    assert np.isnan(y[2]) and np.all(y[[0, 1, 3, 4, 5]] == 6.0)
    with pytest.raises(ValueError):
        device_matvec(A, torch.ones(5, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        device_matvec(A.float(), torch.ones(6, dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("index_set,dim,degree", [("triangle", 2, 12), ("cube", 2, 9), ("cross", 3, 8)])
def test_varfd_device_bitwise_equals_host_entry(index_set, dim, degree):
    import torch
    from hermipy_b200 import core
    from hermipy_b200.plan import varfd_device
    n = core.iterator_size(dim, degree, index_set=index_set)
    rng = np.random.default_rng(n)
    var = rng.standard_normal((n, n))
    for direction in range(dim):
        want = core.varfd(dim, degree, direction, var, index_set=index_set)
        got = varfd_device(dim, degree, direction, torch.from_numpy(var).cuda(), index_set=index_set)
        assert np.array_equal(got.cpu().numpy(), want)
    # a row block with a padded pitch
    store = torch.zeros((7, n + 3), dtype=torch.float64, device="cuda")
    store[:, :n] = torch.from_numpy(var[5:12]).cuda()
    got = varfd_device(dim, degree, 0, store[:, :n], index_set=index_set)
    assert np.array_equal(got.cpu().numpy(), core.varfd(dim, degree, 0, var, index_set=index_set)[5:12])


def _generator_2d(hm):
    x, y, f = hm.x, hm.y, hm.f
    V = x ** 4 / 4 - x ** 2 / 2 + y ** 2 / 2 + x * y / 2
    fxy = f(x, y)
    op = sym.diff(sym.diff(V, x) * fxy + sym.diff(fxy, x), x) + sym.diff(sym.diff(V, y) * fxy + sym.diff(fxy, y), y)
    return op, V


def test_resident_varf_equals_host_varf():
    import hermipy_b200 as hm
    quad = hm.Quad.gauss_hermite(40, dim=2, cov=[[0.5, 0], [0, 0.7]])
    fn = hm.x ** 2 * hm.y + sym.cos(hm.x)
    for index_set, degree in (("triangle", 20), ("cube", 12)):
        host = quad.varf(fn, degree, index_set=index_set, tensorize=False, sparse=False)
        res = quad.varf(fn, degree, index_set=index_set, resident=True)
        assert res.degree == degree and res.matrix.block.is_cuda
        assert np.array_equal(res.matrix.to_host(), host.matrix)
        assert res.to_varf() == host
        d_host = quad.varfd(fn, degree, [0, 1], index_set=index_set, tensorize=False, sparse=False)
        d_res = quad.varfd(fn, degree, [0, 1], index_set=index_set, resident=True)
        np.testing.assert_allclose(d_res.matrix.to_host(), d_host.matrix, rtol=0, atol=1e-15 * np.abs(d_host.matrix).max())


def test_resident_operator_call_solve_eigs():
    import hermipy_b200 as hm
    op, V = _generator_2d(hm)
    degree = 24
    quad = hm.Quad.gauss_hermite(60, dim=2, cov=[[0.4, 0], [0, 0.8]], factor=sym.exp(-V / 2))
    host = quad.discretize_op(op, degree, sparse=False, index_set="triangle")
    res = quad.discretize_op(op, degree, index_set="triangle", resident=True)
    A = host.matrix
    scale = np.abs(A).max()
    np.testing.assert_allclose(res.matrix.to_host(), A, rtol=0, atol=1e-13 * scale)
    # algebra
    np.testing.assert_allclose((res * 2 - res).matrix.to_host(), A, rtol=0, atol=1e-13 * scale)
    # __call__
    rng = np.random.default_rng(0)
    u = hm.Series(rng.standard_normal(A.shape[0]) / (1 + np.arange(A.shape[0])), quad.position, factor=quad.factor)
    np.testing.assert_allclose(res(u).coeffs, host(u).coeffs, rtol=0, atol=1e-13 * scale)
    with pytest.raises(ValueError):
        res(hm.Series(np.ones(3), hm.Position(dim=1)))
    # eigs: the leading (least negative) eigenvalues of the generator; 0 belongs to the invariant density
    # (regular-mode ARPACK needs a generous Krylov space here: the spectrum reaches -8e3 and with the default
    # ncv it returns converged but non-extremal pairs for some start vectors, for ndarray and operator alike)
    v0 = np.random.default_rng(1).standard_normal(A.shape[0])
    want = np.sort(np.linalg.eigvals(A).real)[-4:]
    got, vecs = res.eigs(k=4, which='LR', ncv=80, v0=v0)
    np.testing.assert_allclose(np.sort(got.real), want, atol=1e-8)
    assert abs(np.sort(got.real)[-1]) < 1e-2          # truncation at degree 24 leaves 5.8e-4
    lead = vecs[int(np.argmax(got.real))]
    resid = res(lead).coeffs - got[int(np.argmax(got.real))].real * lead.coeffs
    assert np.linalg.norm(resid) < 1e-8 * max(1.0, np.linalg.norm(lead.coeffs))
    # shift-invert through the device LU
    got_si, _ = res.eigs(k=3, sigma=0.5)
    want_si, _ = host.eigs(k=3, sigma=0.5)
    np.testing.assert_allclose(np.sort(got_si.real), np.sort(want_si.real), atol=1e-8)
    # solve (A - I) v = u: GMRES on the device product and the device LU, against numpy
    shifted_host = host - hm.Varf(np.eye(A.shape[0]), quad.position, factor=quad.factor)
    from hermipy_b200.resident import ResidentVarf
    import torch
    shifted = res - ResidentVarf(torch.eye(A.shape[0], dtype=torch.float64, device="cuda"), quad.position,
                                 factor=quad.factor)
    want_v = shifted_host.solve(u).coeffs
    v_lu = shifted.solve(u, use_gmres=False).coeffs
    np.testing.assert_allclose(v_lu, want_v, rtol=0, atol=1e-10 * np.abs(want_v).max())
    v_gm = shifted.solve(u, rtol=1e-12, restart=200, maxiter=50, quiet=True).coeffs
    np.testing.assert_allclose(v_gm, want_v, rtol=0, atol=1e-8 * np.abs(want_v).max())
    # bordered solve: remove the (given) null vector of the singular generator
    null = lead * (1.0 / np.sqrt(float(lead * lead)))
    rhs = u - null * float(u * null)
    want_b = host.solve(rhs, remove_vec=null, quiet=True).coeffs
    got_b = res.solve(rhs, use_gmres=False, remove_vec=null, quiet=True).coeffs
    np.testing.assert_allclose(got_b, want_b, rtol=0, atol=1e-8 * np.abs(want_b).max())


def test_resident_operator_against_the_reference_package_golden():
    """The resident consumers against the ORACLE: the matrix of the McKean-Vlasov-type generator of tests/api_cases.py
    (case_2d, '2d/operator') as the reference's unmodified Python package produced it (tests/golden/api.npz).  The
    resident matrix, its product, the device-side GMRES, the device LU and the leading eigenvalues are compared with
    NumPy on that golden matrix."""
    import os
    import torch
    import hermipy_b200 as hm
    import api_cases
    from hermipy_b200.resident import ResidentVarf
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api.npz"))["2d/operator"]
    x, y, f = sym.symbols('x y', real=True) + (sym.Function('f'),)
    quad = api_cases.quad_gh(hm, [44, 44], mean=[.2, -.1], cov=[[.8, 0], [0, 1.3]])
    op = sym.diff(f(x, y), x, x) / 2 + sym.diff((x ** 3 - x + y / 2) * f(x, y), x) \
        + sym.diff(f(x, y), y, y) + sym.diff(y * f(x, y), y)
    res = quad.discretize_op(op, 10, index_set="triangle", resident=True)
    A = res.matrix.to_host()
    assert A.shape == golden.shape
    assert np.linalg.norm(A - golden) <= 1e-12 * np.linalg.norm(golden)
    rng = np.random.default_rng(12)
    n = golden.shape[0]
    u = hm.Series(rng.standard_normal(n), quad.position, factor=quad.factor)
    np.testing.assert_allclose(res(u).coeffs, golden @ u.coeffs, rtol=0, atol=1e-12 * np.abs(golden).max() * np.abs(u.coeffs).sum())
    shift = 3.0
    shifted = res - ResidentVarf(shift * torch.eye(n, dtype=torch.float64, device="cuda"), quad.position, factor=quad.factor)
    want = np.linalg.solve(golden - shift * np.eye(n), u.coeffs)
    got_gm = shifted.solve(u, rtol=1e-13, restart=n, maxiter=5, quiet=True).coeffs
    got_lu = shifted.solve(u, use_gmres=False).coeffs
    assert np.linalg.norm(got_gm - want) <= 1e-9 * np.linalg.norm(want)
    assert np.linalg.norm(got_lu - want) <= 1e-10 * np.linalg.norm(want)
    lead = np.sort(np.linalg.eigvals(golden).real)[-3:]
    vals, _ = res.eigs(k=3, which='LR', ncv=min(n - 1, 50), v0=rng.standard_normal(n))
    np.testing.assert_allclose(np.sort(vals.real), lead, atol=1e-8 * max(1.0, np.abs(lead).max()))
