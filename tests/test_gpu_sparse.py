"""GPU parity tests (-m gpu) of the sparse paths, through the C ABI: stored entries in, stored entries out, no dense
intermediate.  Reference: the spmat instantiations of varf (varf.cpp:458-460), tensorize (tensorize.cpp:197-283),
project (project.cpp:78-109), varfd (varf.cpp:290-376); exact zeros are dropped (sparse.cpp:52)."""
import numpy as np
import pytest
import scipy.sparse as ss

from conftest import gauss_hermite

pytestmark = pytest.mark.gpu
SETS = ["triangle", "cross", "cross_nc", "cube", "rectangle"]


def _same_csr(a, b):
    """identical structure and bits (both canonical: sorted columns, no explicit zeros)"""
    assert a.shape == b.shape
    a, b = a.copy(), b.copy()
    b.eliminate_zeros()
    b.sort_indices()
    assert a.has_sorted_indices or a.nnz == 0
    assert (a.data != 0).all()
    assert np.array_equal(a.indptr, b.indptr)
    assert np.array_equal(a.indices, b.indices)
    assert np.array_equal(a.data, b.data)


def _poly2(X, Y):
    return X ** 4 / 4 - X ** 2 / 2 + Y ** 2 / 2 + X * Y / 2 + 0.3


@pytest.mark.parametrize("index_set", SETS)
def test_varf_sparse_is_the_dense_result_compressed(core, ref, index_set):
    """polynomial f: the band of retained triple products goes straight to CSR; same entries, bit for bit, as the dense
    assembly, and the reference's matrix to 1e-12 with its zero pattern"""
    x, w = gauss_hermite(40)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = _poly2(X, Y)
    deg = 14
    dense = core.varf(deg, f, [x, x], [w, w], index_set=index_set)
    sp = core.varf(deg, f, [x, x], [w, w], sparse=True, index_set=index_set)
    assert isinstance(sp, ss.csr_matrix)
    _same_csr(sp, ss.csr_matrix(dense))
    want = ref.varf(deg, f, [x, x], [w, w], index_set=index_set)
    assert np.linalg.norm(sp.toarray() - want) <= 1e-12 * np.linalg.norm(want)
    assert np.array_equal(sp.toarray() != 0, want != 0)


def test_varf_sparse_3d_and_1d(core, ref):
    x, w = gauss_hermite(24)
    f1 = x ** 3 - 2 * x + 1
    _same_csr(core.varf(10, f1, [x], [w], sparse=True), ss.csr_matrix(core.varf(10, f1, [x], [w])))
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    f3 = X * Y + Z ** 2 - 0.5
    d3 = core.varf(5, f3, [x] * 3, [w] * 3, index_set="triangle")
    _same_csr(core.varf(5, f3, [x] * 3, [w] * 3, sparse=True, index_set="triangle"), ss.csr_matrix(d3))


def test_varf_sparse_dense_f_falls_back_to_compaction(core):
    """a non-polynomial f keeps many alpha: no narrow band, the dense device matrix is compacted"""
    x, w = gauss_hermite(30)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = np.exp(-(X ** 2 + Y ** 2) / 7) * (1 + X)
    dense = core.varf(8, f, [x, x], [w, w], index_set="cube")
    _same_csr(core.varf(8, f, [x, x], [w, w], sparse=True, index_set="cube"), ss.csr_matrix(dense))


def test_varf_sparse_c2_size_needs_no_dense_buffer(core):
    """C2 (degree 200 cube, 40 401^2 = 13 GB dense): the sparse assembly of a degree-4 polynomial touches a few MB"""
    import torch
    x, w = gauss_hermite(401)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    sp = core.varf(200, f, [x, x], [w, w], sparse=True, index_set="cube")
    torch.cuda.synchronize()
    used = free0 - torch.cuda.mem_get_info()[0]
    assert sp.shape == (40401, 40401) and 40401 * 5 < sp.nnz < 40401 * 81
    assert used < 4e9, "the sparse varf allocated %.1f GB on the device" % (used * 1e-9)
    # <V phi_m phi_n> in closed form through powers of the Jacobi matrix (x phi_k = sqrt(k+1) phi_{k+1} + sqrt(k) phi_{k-1})
    J = np.diag(np.sqrt(np.arange(1.0, 206)), 1)
    J = J + J.T
    Jp = [np.linalg.matrix_power(J, p)[:201, :201] for p in range(5)]
    idx = core.iterator_list_indices(2, 200, "cube")
    rng = np.random.default_rng(3)
    rows = rng.integers(0, 40401, 300)
    sub = sp[rows].toarray()
    r0, r1, c0, c1 = idx[rows, 0], idx[rows, 1], idx[:, 0], idx[:, 1]
    pw = lambda a, b: Jp[a][r0[:, None], c0[None, :]] * Jp[b][r1[:, None], c1[None, :]]
    exact = pw(4, 0) / 4 - pw(2, 0) / 2 + pw(0, 4) / 4 - pw(0, 2) / 2 + pw(1, 1) / 2
    assert np.linalg.norm(sub - exact) <= 1e-11 * np.linalg.norm(exact)


@pytest.mark.parametrize("index_set", SETS)
def test_tensorize_sparse_factors(core, index_set):
    rng = np.random.default_rng(1)
    deg = 9
    n1 = core.iterator_size(1, deg, index_set)
    A = ss.random(n1, n1, density=0.3, random_state=2, format="csr")
    B = ss.random(n1, n1, density=0.25, random_state=3, format="csr")
    dense = core.tensorize([A.toarray(), B.toarray()], index_set=index_set)
    sp = core.tensorize([A, B], sparse=True, index_set=index_set)
    assert isinstance(sp, ss.csr_matrix)
    _same_csr(sp, ss.csr_matrix(dense))
    # dense factors, sparse result; sparse factors, dense result
    _same_csr(core.tensorize([A.toarray(), B.toarray()], sparse=True, index_set=index_set), ss.csr_matrix(dense))
    assert np.array_equal(core.tensorize([A, B], index_set=index_set), dense)
    # three factors, one of them over two directions, directions out of order
    n2 = core.iterator_size(2, deg, index_set)
    C2 = ss.random(n2, n2, density=0.05, random_state=4, format="csr")
    inp = {frozenset({1}): A, frozenset({0, 2}): C2}
    dense3 = core.tensorize({k: v.toarray() for k, v in inp.items()}, index_set=index_set)
    _same_csr(core.tensorize(inp, sparse=True, index_set=index_set), ss.csr_matrix(dense3))
    # one factor and identities elsewhere
    dense_id = core.tensorize(A.toarray(), dim=3, direction=1, index_set=index_set)
    _same_csr(core.tensorize(A, dim=3, direction=1, sparse=True, index_set=index_set), ss.csr_matrix(dense_id))


@pytest.mark.parametrize("index_set", SETS)
def test_project_and_varfd_sparse(core, index_set):
    deg, dim = 6, 3
    n = core.iterator_size(dim, deg, index_set)
    M = ss.random(n, n, density=0.08, random_state=5, format="csr")
    for dirs in ([0], [2], [0, 2], [1, 2]):
        _same_csr(core.project(M, dim, dirs, index_set=index_set),
                  ss.csr_matrix(core.project(M.toarray(), dim, dirs, index_set=index_set)))
    for direction in range(dim):
        _same_csr(core.varfd(dim, deg, direction, M, index_set=index_set),
                  ss.csr_matrix(core.varfd(dim, deg, direction, M.toarray(), index_set=index_set)))
    # Fourier direction (even degree)
    _same_csr(core.varfd(dim, deg, 1, M, do_fourier=1, index_set=index_set),
              ss.csr_matrix(core.varfd(dim, deg, 1, M.toarray(), do_fourier=1, index_set=index_set)))


def test_sparse_edge_cases(core):
    n = core.iterator_size(2, 4, "triangle")
    empty = ss.csr_matrix((n, n))
    assert core.project(empty, 2, [0]).nnz == 0
    assert core.varfd(2, 4, 0, empty).nnz == 0
    n1 = core.iterator_size(1, 4, "triangle")
    assert core.tensorize([ss.csr_matrix((n1, n1)), ss.identity(n1, format="csr")], sparse=True).nnz == 0
    eye = core.tensorize([ss.identity(n1, format="csr")] * 2, sparse=True)
    _same_csr(eye, ss.identity(n, format="csr"))


def test_plan_varf_sparse_row_blocks(core):
    """row-sharded sparse varf: the owned rows as CSR equal the rows of the full sparse matrix"""
    from hermipy_b200.plan import Plan
    x, w = gauss_hermite(48)
    X, Y = np.meshgrid(x, x, indexing="ij")
    plan = Plan(20, [x, x], [w, w], index_set="triangle")
    f = plan.pack_grid(_poly2(X, Y)[None])[0]
    full = plan.varf_sparse(f)
    _same_csr(full, ss.csr_matrix(plan.varf(f, "reference").cpu().numpy()))
    n = plan.n_polys
    for lo, hi in ((0, n // 3), (n // 3, n - 7), (n - 7, n)):
        _same_csr(plan.varf_sparse(f, lo, hi), full[lo:hi])
