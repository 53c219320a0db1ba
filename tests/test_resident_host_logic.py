"""CPU tests (-m "not gpu") of the host-side logic of hermipy_b200.resident (reference: hermipy/varf.py:127-237):
operator algebra, the bordered system of `solve(remove_vec=...)`, the ARPACK / GMRES drivers.  The matrix-vector
product is a CPU stand-in injected through `ResidentMatrix(matvec=...)`; the device product is covered by
tests/test_gpu_resident.py."""
import numpy as np
import pytest
import torch

import oracle_backend

from hermipy_b200 import Position, Series
from hermipy_b200.resident import ResidentMatrix, ResidentVarf


def cpu_matvec(A, x, out=None):
    res = torch.mv(A, x)
    if out is None:
        return res
    out.copy_(res)
    return out


@pytest.fixture()
def operator():
    rng = np.random.default_rng(4)
    n = 28                                         # triangle set, dim 2, degree 6
    A = rng.standard_normal((n, n)) * 0.3 - 3.0 * np.eye(n)
    position = Position(dim=2)
    varf = ResidentVarf(ResidentMatrix(torch.from_numpy(A.copy()), matvec=cpu_matvec), position)
    return A, position, varf


def test_degree_indices_and_call(operator):
    A, position, varf = operator
    assert varf.degree == 6 and len(varf.multi_indices()) == 28
    u = Series(np.arange(28.0), position)
    np.testing.assert_allclose(varf(u).coeffs, A @ np.arange(28.0), rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        varf(Series(np.ones(7), Position(dim=1)))


def test_algebra(operator):
    A, position, varf = operator
    np.testing.assert_allclose((varf * 2 - varf).matrix.to_host(), A, rtol=0, atol=1e-15)
    np.testing.assert_allclose((-varf + varf * 3.0).matrix.to_host(), 2 * A, rtol=0, atol=1e-14)
    np.testing.assert_allclose((0 + varf).matrix.to_host(), A)
    np.testing.assert_allclose((varf + 1.5).matrix.to_host(), A + 1.5)
    with pytest.raises(ValueError):
        varf * varf
    with pytest.raises(ValueError):
        varf + ResidentVarf(varf.matrix, position, index_set="cube" if False else "triangle", factor=2)
    host = varf.to_varf()
    np.testing.assert_allclose(host.matrix, A)


def test_solve_gmres_and_direct(operator):
    A, position, varf = operator
    b = Series(np.linspace(-1, 1, 28), position)
    want = np.linalg.solve(A, b.coeffs)
    got = varf.solve(b, rtol=1e-13, restart=28, quiet=True).coeffs
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-10)
    got_lu = varf.solve(b, use_gmres=False).coeffs
    np.testing.assert_allclose(got_lu, want, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        varf.solve(b, preconditioner="ilu")


def test_bordered_solve_matches_explicit_system(operator):
    A, position, varf = operator
    rng = np.random.default_rng(5)
    v = Series(rng.standard_normal(28), position)
    b = Series(rng.standard_normal(28), position)
    vn = v.coeffs / np.sqrt(v.coeffs @ v.coeffs)
    big = np.block([[A, vn[:, None]], [vn[None, :], np.zeros((1, 1))]])          # varf.py:200-206
    want = np.linalg.solve(big, np.append(b.coeffs, 0))[:-1]
    with oracle_backend.oracle_core():            # Series * Series (the normalisation of remove_vec) is core.inner
        got_lu = varf.solve(b, use_gmres=False, remove_vec=v, quiet=True).coeffs
        got_gm = varf.solve(b, remove_vec=v, rtol=1e-13, restart=29, quiet=True).coeffs
    np.testing.assert_allclose(got_lu, want, rtol=0, atol=1e-11)
    np.testing.assert_allclose(got_gm, want, rtol=0, atol=1e-9)


def test_eigs_regular_and_shift_invert(operator):
    A, position, varf = operator
    true = np.linalg.eigvals(A)
    values, vectors = varf.eigs(k=3, which="LM", v0=np.ones(28))
    for lam, vec in zip(values, vectors):
        assert np.min(np.abs(true - lam)) < 1e-8
        resid = A @ vec.coeffs - lam * vec.coeffs
        assert np.linalg.norm(resid) < 1e-8 * np.linalg.norm(vec.coeffs)
    sigma = -3.0
    values, _ = varf.eigs(k=2, sigma=sigma, v0=np.ones(28))
    nearest = true[np.argsort(np.abs(true - sigma))[:2]]
    assert all(np.min(np.abs(nearest - lam)) < 1e-8 for lam in values)


def test_device_gmres_matches_a_direct_solve_and_scipy_semantics():
    """hermipy_b200.resident.device_gmres (the Krylov driver of ResidentVarf.solve) on CPU tensors with a torch.mv
    stand-in for the device product: converges to the direct solution, honours rtol / restart / maxiter like
    scipy.sparse.linalg.gmres (info = 0 on success, > 0 otherwise), exact for a rank-one update of the identity."""
    import scipy.sparse.linalg as spla
    import torch
    from hermipy_b200.resident import device_gmres
    rng = np.random.default_rng(4)
    n = 120
    A = np.eye(n) * 4 + rng.standard_normal((n, n)) / np.sqrt(n)
    b = rng.standard_normal(n)
    At, bt = torch.from_numpy(A), torch.from_numpy(b)
    want = np.linalg.solve(A, b)
    x, info = device_gmres(lambda v: torch.mv(At, v), bt, rtol=1e-12, restart=40, maxiter=20)
    assert info == 0 and np.linalg.norm(x.numpy() - want) <= 1e-10 * np.linalg.norm(want)
    # same tolerance semantics as SciPy: residual <= rtol * ||b||
    xs, info_s = spla.gmres(A, b, rtol=1e-6, restart=20)
    xd, info_d = device_gmres(lambda v: torch.mv(At, v), bt, rtol=1e-6, restart=20)
    assert info_s == 0 and info_d == 0
    assert np.linalg.norm(A @ xd.numpy() - b) <= 1e-6 * np.linalg.norm(b) * (1 + 1e-6)
    assert np.linalg.norm(xd.numpy() - xs) <= 1e-4 * np.linalg.norm(xs)
    # too few iterations: reported, not hidden
    _, info_short = device_gmres(lambda v: torch.mv(At, v), bt, rtol=1e-14, restart=3, maxiter=1)
    assert info_short > 0
    # happy breakdown: A = I + u u^T has a two-dimensional Krylov space
    u = torch.from_numpy(rng.standard_normal(n))
    x2, info2 = device_gmres(lambda v: v + u * (u @ v), bt, rtol=1e-13)
    assert info2 == 0
    assert np.linalg.norm((x2 + u * (u @ x2) - bt).numpy()) <= 1e-12 * np.linalg.norm(b)
    # start vector
    x3, info3 = device_gmres(lambda v: torch.mv(At, v), bt, x0=torch.from_numpy(want + 1e-3), rtol=1e-12, restart=40)
    assert info3 == 0 and np.linalg.norm(x3.numpy() - want) <= 1e-10 * np.linalg.norm(want)


def test_device_eigs_against_numpy_and_arpack():
    """hermipy_b200.resident.device_eigs (Krylov-Schur, the driver of ResidentVarf.eigs) on CPU tensors with a torch.mv
    stand-in for the device product: wanted eigenvalues for every `which`, residuals of the returned pairs, agreement
    with scipy.sparse.linalg.eigs (ARPACK), complex pairs kept together."""
    import scipy.sparse.linalg as spla
    import torch
    from hermipy_b200.resident import device_eigs
    rng = np.random.default_rng(0)
    n = 200
    A = -np.diag(np.arange(n, dtype=float)) * 0.5 + rng.standard_normal((n, n)) * 0.3
    At = torch.from_numpy(A)
    ev = np.linalg.eigvals(A)
    keys = {"LR": -ev.real, "LM": -np.abs(ev), "SR": ev.real, "SM": np.abs(ev)}
    for which, k in (("LR", 4), ("LM", 3), ("SR", 3), ("SM", 2)):
        vals, vecs = device_eigs(lambda v: torch.mv(At, v), n, k=k, which=which, ncv=50, tol=1e-10)
        want = ev[np.argsort(keys[which], kind="stable")[:k]]
        # the k-th and (k+1)-th wanted values may be a conjugate pair: compare the set of |.|-sorted values loosely
        assert np.allclose(np.sort(np.abs(vals)), np.sort(np.abs(want)), atol=1e-7), (which, vals, want)
        V = vecs.numpy()
        assert np.linalg.norm(A @ V - V * np.asarray(vals)[None, :]) <= 1e-8 * np.linalg.norm(A)
    ar = spla.eigs(A, k=4, which="LR", ncv=50, v0=np.ones(n))[0]
    dv, _ = device_eigs(lambda v: torch.mv(At, v), n, k=4, which="LR", ncv=50, v0=torch.ones(n, dtype=torch.float64))
    assert np.allclose(np.sort_complex(np.asarray(dv, dtype=complex)), np.sort_complex(ar), atol=1e-8)
    # symmetric matrix: real eigenpairs come back real
    S = (A + A.T) / 2
    St = torch.from_numpy(S)
    sv, sV = device_eigs(lambda v: torch.mv(St, v), n, k=3, which="LR", ncv=40)
    assert sv.dtype == np.float64 and not sV.is_complex()
    assert np.allclose(np.sort(sv), np.sort(np.linalg.eigvalsh(S))[-3:], atol=1e-9)
