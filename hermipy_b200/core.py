"""Drop-in counterpart of the reference's hermipy/core.py (hermipy/core.py:125-348): the same function
names, argument meaning and return types, but every numerical call goes through the C ABI of
libhermite_b200.so (hand-written sm_100a CUDA) instead of the Boost.Python module `hermite_cpp`.

NumPy arrays cross the boundary as contiguous float64 buffers (no per-element boxing as in
core.py:29-66 of the reference).  Invalid input raises an exception where the reference would
print a message and exit() the interpreter.
"""
import ctypes as C
import threading

import numpy as np
import scipy.sparse as ss

from . import _lib
from ._lib import c_dp, c_ip, c_llp, c_up, check, lib

VARF_MODE = {"reference": _lib.VARF_REFERENCE, "gram": _lib.VARF_GRAM}


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _grid(nodes, weights):
    nodes = [_d(n).ravel() for n in nodes]
    weights = [_d(w).ravel() for w in weights]
    n_pts = _i([len(n) for n in nodes])
    return n_pts, _d(np.concatenate(nodes)), _d(np.concatenate(weights))


def _fourier(do_fourier, dim):
    return _i([0] * dim if do_fourier is None else do_fourier)


# ---- iterator functions (hermipy/core.py:283-346) --------------------------------------------------

_SETS = ("triangle", "cross", "cross_nc", "cube", "rectangle")


def _check_set(index_set):
    if index_set not in _SETS:
        raise ValueError("Unknown index set")


def iterator_get_degree(dim, n_polys, index_set="triangle"):
    _check_set(index_set)
    out = C.c_int(0)
    check(lib().hb_index_get_degree(index_set.encode(), int(dim), int(n_polys), C.byref(out)))
    return out.value


def iterator_index(mult_ind, index_set="triangle"):
    m = np.ascontiguousarray(mult_ind, dtype=np.uint32)
    fn = {"triangle": lib().hb_triangle_index, "cube": lib().hb_cube_index}
    if index_set not in fn:
        raise ValueError("Unknown index set")
    out = C.c_uint32(0)
    check(fn[index_set](m.ctypes.data_as(c_up), len(m), C.byref(out)))
    return int(out.value)


def iterator_positions(mult_inds, degree, index_set="triangle"):
    """Positions of many multi-indices (count x dim) in the enumeration of the set; -1 where absent."""
    _check_set(index_set)
    m = np.ascontiguousarray(mult_inds, dtype=np.uint32)
    if m.ndim != 2:
        raise ValueError("mult_inds must be a (count, dim) array")
    out = np.empty(m.shape[0], dtype=np.int32)
    check(lib().hb_index_positions(index_set.encode(), m.shape[1], int(degree), m.ctypes.data_as(c_up), m.shape[0],
                                   out.ctypes.data_as(c_ip)))
    return out


def iterator_size(dim, degree, index_set="triangle"):
    _check_set(index_set)
    out = C.c_longlong(0)
    check(lib().hb_index_size(index_set.encode(), int(dim), int(degree), C.byref(out)))
    return int(out.value)


def iterator_list_indices(dim, degree, index_set="triangle"):
    n = iterator_size(dim, degree, index_set)
    out = np.zeros((n, dim), dtype=np.uint32)
    check(lib().hb_index_list(index_set.encode(), int(dim), int(degree), out.ctypes.data_as(c_up), out.size))
    return np.asarray(out, dtype=int)


# ---- wrapper functions (hermipy/core.py:125-281) ---------------------------------------------------

def hermite_eval(x, degree):
    """Normalised Hermite polynomials He_k(x)/sqrt(k!) for k <= degree at every x (hermite.cpp:29-41)."""
    x = _d(np.atleast_1d(x))
    out = np.zeros((len(x), degree + 1))
    check(lib().hb_hermite_eval(_dp(x), len(x), int(degree), _dp(out)))
    return out


def discretize(function, nodes, translation, dilation):
    """hermipy/core.py:128-130: `function` is the C expression in v[i] produced by Function.ccode()
    (function.py:118-123); evaluated on the GPU at translation + dilation @ node for every grid node."""
    nodes = [_d(n).ravel() for n in nodes]
    n_pts = _i([len(n) for n in nodes])
    dim = len(nodes)
    nd = _d(np.concatenate(nodes))
    tr = _d(translation).ravel()
    di = _d(dilation).reshape(dim, dim)
    out = np.zeros(int(np.prod(n_pts)))
    check(lib().hb_discretize(str(function).encode(), dim, _ip(n_pts), _dp(nd), _dp(tr), _dp(di), _dp(out), out.size))
    return out


def integrate(fgrid, nodes, weights, do_fourier=None):
    n_pts, nd, wt = _grid(nodes, weights)
    f = _d(fgrid).ravel()
    out = C.c_double(0)
    check(lib().hb_integrate(len(n_pts), _ip(n_pts), _dp(f), _dp(nd), _dp(wt), _ip(_fourier(do_fourier, len(n_pts))),
                             C.byref(out)))
    return out.value


def transform(degree, fgrid, nodes, weights, forward, do_fourier=None, index_set="triangle"):
    n_pts, nd, wt = _grid(nodes, weights)
    dim = len(n_pts)
    degree = int(degree)
    f = _d(fgrid).ravel()
    n_grid = int(np.prod(n_pts))
    n_polys = iterator_size(dim, degree, index_set)
    n_in, n_out = (n_grid, n_polys) if forward else (n_polys, n_grid)
    if len(f) != n_in:
        raise ValueError("transform: input has %d entries, expected %d" % (len(f), n_in))
    out = np.zeros(n_out)
    check(lib().hb_transform(dim, _ip(n_pts), degree, _dp(f), _dp(nd), _dp(wt), _ip(_fourier(do_fourier, dim)),
                             int(bool(forward)), index_set.encode(), _dp(out), n_out))
    return out


_SIZES = {}


def transform_batch(degree, fgrids, nodes, weights, forward, do_fourier=None, index_set="triangle", out=None):
    """`batch` independent transforms in one call; fgrids has shape (batch, n_in).  No reference equivalent
    (the reference loops over hm.Quad.transform in Python).  `out` may be a preallocated (batch, n_out)
    float64 array (e.g. page-locked memory) to receive the result."""
    n_pts, nd, wt = _grid(nodes, weights)
    dim = len(n_pts)
    f = _d(fgrids)
    batch = f.shape[0]
    n_grid = int(np.prod(n_pts))
    key = (dim, int(degree), index_set)
    n_polys = _SIZES.get(key)
    if n_polys is None:                       # the decorated iterator_size costs 20 us per call
        n_polys = _SIZES[key] = iterator_size(dim, int(degree), index_set)
    n_in, n_out = (n_grid, n_polys) if forward else (n_polys, n_grid)
    if batch == 0:
        return np.zeros((0, n_out))
    f = f.reshape(batch, -1)
    if f.shape[1] != n_in:
        raise ValueError("transform_batch: rows have %d entries, expected %d" % (f.shape[1], n_in))
    if out is None:
        out = np.zeros((batch, n_out))
    elif out.shape != (batch, n_out) or out.dtype != np.float64 or not out.flags.c_contiguous:
        raise ValueError("transform_batch: out must be a C-contiguous float64 array of shape (%d, %d)" % (batch, n_out))
    check(lib().hb_transform_batch(dim, _ip(n_pts), int(degree), batch, _dp(f), _dp(nd), _dp(wt),
                                   _ip(_fourier(do_fourier, dim)), int(bool(forward)), index_set.encode(),
                                   _dp(out), n_out))
    return out


def transform_weighted(degree, function, factor, nodes, weights, translation, dilation, do_fourier=None,
                       index_set="triangle"):
    """Coefficients of function / factor (hermipy/quad.py:287-306) in ONE device pass: `function` is a grid array
    (one function, or a batch of shape (batch, n_points)) or a C expression in v[i]; `factor` a C expression or None;
    expressions are evaluated on the device at translation + dilation @ node, the division
    f / (factor == 0 ? 1e-300 : factor) is a device kernel, and only the inputs and the coefficients cross PCIe."""
    n_pts, nd, wt = _grid(nodes, weights)
    dim = len(n_pts)
    n_grid = int(np.prod(n_pts))
    n_polys = iterator_size(dim, int(degree), index_set)
    tr, di = _d(translation).ravel(), _d(dilation).reshape(dim, dim)
    fac = None if factor is None else str(factor).encode()
    if isinstance(function, str):
        f_arr, f_body, batch, single = None, function.encode(), 1, True
    else:
        f_arr = _d(function)
        single = f_arr.size == n_grid and (f_arr.ndim != 2 or f_arr.shape[0] != 1 or n_grid == 1)
        f_arr = f_arr.reshape(-1, n_grid)
        f_body, batch = None, f_arr.shape[0]
    out = np.zeros((batch, n_polys))
    check(lib().hb_transform_weighted(dim, _ip(n_pts), int(degree), batch, None if f_arr is None else _dp(f_arr), f_body,
                                      fac, _dp(nd), _dp(wt), _dp(tr), _dp(di), _ip(_fourier(do_fourier, dim)),
                                      index_set.encode(), _dp(out), n_polys))
    return out[0] if single else out


def eval_weighted(degree, coeffs, factor, eval_nodes, nodes, weights, translation, dilation, do_fourier=None,
                  index_set="triangle"):
    """Values of a series on `eval_nodes` times `factor` (a C expression, or None) evaluated on `nodes` at
    translation + dilation @ node (hermipy/quad.py:316-328), without leaving the device in between."""
    n_pts, nd, wt = _grid(nodes, weights)
    dim = len(n_pts)
    ev = _d(np.concatenate([_d(n).ravel() for n in eval_nodes]))
    c = _d(coeffs).ravel()
    tr, di = _d(translation).ravel(), _d(dilation).reshape(dim, dim)
    out = np.zeros(int(np.prod(n_pts)))
    check(lib().hb_eval_weighted(dim, _ip(n_pts), int(degree), _dp(c), len(c), _dp(ev), _dp(nd), _dp(wt),
                                 None if factor is None else str(factor).encode(), _dp(tr), _dp(di),
                                 _ip(_fourier(do_fourier, dim)), index_set.encode(), _dp(out), out.size))
    return out


def varf_function(degree, function, nodes, weights, translation, dilation, sparse=False, do_fourier=None,
                  index_set="triangle", mode="reference"):
    """core.varf of a function given as a C expression in v[i]: evaluated on the device grid, never on the host."""
    n_pts, nd, wt = _grid(nodes, weights)
    dim = len(n_pts)
    degree = int(degree)
    n = iterator_size(dim, degree, index_set)
    tr, di = _d(translation).ravel(), _d(dilation).reshape(dim, dim)
    four = _fourier(do_fourier, dim)
    body = str(function).encode()
    if not sparse:
        out = np.zeros((n, n))
        check(lib().hb_varf_function(dim, _ip(n_pts), degree, body, _dp(tr), _dp(di), _dp(nd), _dp(wt), _ip(four),
                                     index_set.encode(), VARF_MODE[mode], _dp(out), n))
        return out
    with _SP_LOCK:
        nnz = C.c_longlong(0)
        check(lib().hb_varf_function_sp(dim, _ip(n_pts), degree, body, _dp(tr), _dp(di), _dp(nd), _dp(wt), _ip(four),
                                        index_set.encode(), VARF_MODE[mode], C.byref(nnz)))
        return _fetch_csr(n, n, nnz.value)


def triple_products(degree):
    out = np.zeros((2 * degree + 1, degree + 1, degree + 1))
    check(lib().hb_triple_products(int(degree), _dp(out)))
    return out


def triple_products_fourier(degree):
    out = np.zeros((2 * degree + 1, degree + 1, degree + 1))
    check(lib().hb_triple_products_fourier(int(degree), _dp(out)))
    return out


def varf(degree, fgrid, nodes, weights, sparse=False, do_fourier=None, index_set="triangle", mode="reference"):
    n_pts, nd, wt = _grid(nodes, weights)
    dim = len(n_pts)
    degree = int(degree)
    f = _d(fgrid).ravel()
    if len(f) != int(np.prod(n_pts)):
        raise ValueError("varf: input has %d entries, expected %d" % (len(f), int(np.prod(n_pts))))
    n = iterator_size(dim, degree, index_set)
    four = _fourier(do_fourier, dim)
    if not sparse:
        out = np.zeros((n, n))
        check(lib().hb_varf(dim, _ip(n_pts), degree, _dp(f), _dp(nd), _dp(wt), _ip(four), index_set.encode(),
                            VARF_MODE[mode], _dp(out), n))
        return out
    with _SP_LOCK:
        nnz = C.c_longlong(0)
        check(lib().hb_varf_sp(dim, _ip(n_pts), degree, _dp(f), _dp(nd), _dp(wt), _ip(four), index_set.encode(),
                               VARF_MODE[mode], C.byref(nnz)))
        return _fetch_csr(n, n, nnz.value)


# The library keeps ONE sparse result per process: a *_sp call and the fetch of its result form one critical section.
_SP_LOCK = threading.Lock()


def _fetch_csr(rows, cols, nnz):
    indptr = np.zeros(rows + 1, dtype=np.int64)
    indices = np.zeros(max(nnz, 1), dtype=np.int32)
    data = np.zeros(max(nnz, 1))
    check(lib().hb_sparse_fetch(indptr.ctypes.data_as(_lib.c_llp), _ip(indices), _dp(data)))
    return ss.csr_matrix((data[:nnz], indices[:nnz], indptr), shape=(rows, cols))


def _csr_arrays(m):
    """(indptr int64, indices int32, data float64) of a matrix in canonical CSR form (format conversion on the host,
    no arithmetic: a dense factor is scanned for its non-zeros as the reference's to_spmat does, sparse.cpp:44-58)"""
    m = m if isinstance(m, ss.csr_matrix) else ss.csr_matrix(np.asarray(m, dtype=np.float64))
    if not m.has_canonical_format:
        m = m.copy()
        m.sum_duplicates()
    return (np.ascontiguousarray(m.indptr, dtype=np.int64), np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=np.float64))


def varfd(dim, degree, direction, var, do_fourier=0, index_set="triangle"):
    if isinstance(var, ss.csr_matrix):      # stored entries in, stored entries out (varf.cpp:290-376)
        indptr, indices, data = _csr_arrays(var)
        n = var.shape[0]
        with _SP_LOCK:
            nnz = C.c_longlong(0)
            check(lib().hb_varfd_sp(int(dim), int(degree), int(direction), indptr.ctypes.data_as(_lib.c_llp), _ip(indices),
                                    _dp(data), n, int(do_fourier), index_set.encode(), C.byref(nnz)))
            return _fetch_csr(n, n, nnz.value)
    v = _d(var)
    out = np.zeros_like(v)
    check(lib().hb_varfd(int(dim), int(degree), int(direction), _dp(v), v.shape[0], int(do_fourier),
                         index_set.encode(), _dp(out)))
    return out


def _tensorize_call(arrays, dirs_list, index_set, sparse):
    k = len(arrays)
    matrices = len(arrays[0].shape) == 2
    lens = _i([a.shape[0] for a in arrays])
    dflat = _i(np.concatenate([np.asarray(d, dtype=np.int32) for d in dirs_list]))
    dlens = _i([len(d) for d in dirs_list])
    dim = int(dlens.sum())
    degree = iterator_get_degree(len(dirs_list[0]), arrays[0].shape[0], index_set)
    n = iterator_size(dim, degree, index_set)
    any_sparse = any(isinstance(a, ss.csr_matrix) for a in arrays)
    if matrices and (sparse or any_sparse):
        # sparse factors or a sparse result: pairs of stored entries, never an n x n buffer (tensorize.cpp:197-283)
        trip = [_csr_arrays(a) for a in arrays]
        p_ptr = (c_llp * k)(*[t[0].ctypes.data_as(c_llp) for t in trip])
        p_ind = (c_ip * k)(*[_ip(t[1]) for t in trip])
        p_dat = (c_dp * k)(*[_dp(t[2]) for t in trip])
        with _SP_LOCK:
            rows, nnz = C.c_longlong(0), C.c_longlong(0)
            check(lib().hb_tensorize_mats_sp(k, p_ptr, p_ind, p_dat, _ip(lens), _ip(dflat), _ip(dlens), index_set.encode(),
                                             C.byref(rows), C.byref(nnz)))
            res = _fetch_csr(n, n, nnz.value)
        return res if sparse else res.toarray()      # sparse factors, dense result requested
    arrays = [_d(a) for a in arrays]
    ptrs = (c_dp * k)(*[_dp(a) for a in arrays])
    if matrices:
        out = np.zeros((n, n))
        check(lib().hb_tensorize_mats(k, ptrs, _ip(lens), _ip(dflat), _ip(dlens), index_set.encode(), _dp(out), n))
        return out
    out = np.zeros(n)
    check(lib().hb_tensorize_vecs(k, ptrs, _ip(lens), _ip(dflat), _ip(dlens), index_set.encode(), _dp(out), n))
    return out


def tensorize(inp, dim=None, direction=None, sparse=False, index_set="triangle"):
    """hermipy/core.py:201-253: `inp` is a dict {frozenset(dirs): array}, a list (one factor per axis), or a
    single array with (dim, direction) (tensorised with identities on the other axes)."""
    if isinstance(inp, dict):
        any_elem = list(inp.values())[0]
        if isinstance(any_elem, (float, int)):
            return np.prod([inp[k] for k in inp])
        assert isinstance(any_elem, (np.ndarray, ss.csr_matrix))
        assert len(any_elem.shape) in (1, 2)
        dirs_list = [sorted(dirs) for dirs in inp.keys()]
        return _tensorize_call(list(inp.values()), dirs_list, index_set, sparse)
    if isinstance(inp, list):
        assert dim is None and direction is None
        if isinstance(inp[0], (float, int)):
            return np.prod(inp)
        return _tensorize_call(inp, [[i] for i in range(len(inp))], index_set, sparse)
    assert dim is not None and direction is not None
    a = inp if isinstance(inp, ss.csr_matrix) else _d(inp)
    # tensorize_vec_id / tensorize_mat_id (tensorize.cpp:145-152, 373-381): identity factors elsewhere
    if len(a.shape) == 1:
        unit = np.zeros(len(a))
        unit[0] = 1.0
    elif sparse or isinstance(a, ss.csr_matrix):
        unit = ss.identity(a.shape[0], format="csr")
    else:
        unit = np.eye(a.shape[0])
    factors = [unit] * dim
    factors[direction] = a
    return _tensorize_call(factors, [[i] for i in range(dim)], index_set, sparse)


def project(inp, dim, directions, index_set="triangle"):
    if isinstance(directions, int):
        directions = [directions]
    sparse = isinstance(inp, ss.csr_matrix)
    if isinstance(inp, list):
        inp = np.asarray(inp, dtype=float)
    dirs = _i(directions)
    degree = iterator_get_degree(dim, inp.shape[0], index_set)
    n = iterator_size(len(dirs), degree, index_set)
    if sparse:      # stored entries in, stored entries out (project.cpp:78-109)
        indptr, indices, data = _csr_arrays(inp)
        with _SP_LOCK:
            rows, nnz = C.c_longlong(0), C.c_longlong(0)
            check(lib().hb_project_mat_sp(indptr.ctypes.data_as(_lib.c_llp), _ip(indices), _dp(data), inp.shape[0], int(dim),
                                          _ip(dirs), len(dirs), index_set.encode(), C.byref(rows), C.byref(nnz)))
            return _fetch_csr(n, n, nnz.value)
    a = _d(inp)
    if a.ndim == 1:
        out = np.zeros(n)
        check(lib().hb_project_vec(_dp(a), len(a), int(dim), _ip(dirs), len(dirs), index_set.encode(), _dp(out), n))
        return out
    out = np.zeros((n, n))
    check(lib().hb_project_mat(_dp(a), a.shape[0], int(dim), _ip(dirs), len(dirs), index_set.encode(), _dp(out), n))
    return out


def inner(s1, s2, d1, d2, index_set="triangle"):
    s1, s2, d1, d2 = _d(s1), _d(s2), _i(d1), _i(d2)
    degree = iterator_get_degree(len(d1), len(s1), index_set)
    n_res = len(set(d1.tolist()) ^ set(d2.tolist()))
    n = iterator_size(n_res, degree, index_set) if n_res > 0 else 1
    out = np.zeros(n)
    check(lib().hb_inner(_dp(s1), len(s1), _dp(s2), len(s2), _ip(d1), len(d1), _ip(d2), len(d2),
                         index_set.encode(), _dp(out), n))
    return out


# ---- decorators of the reference's wrappers (hermipy/core.py:125-348: @cache() @debug() @log_stats()) ----
# Applied here, after the plain definitions, so that the undecorated functions stay reachable as
# `<name>.__wrapped__`.  `cache=` becomes a keyword of every wrapped function (hermipy/cache.py:107-110).
def _decorate():
    from .cache import cache
    from .stats import debug, log_stats
    names = ("discretize", "integrate", "transform", "transform_weighted", "eval_weighted", "varf_function",
             "triple_products", "triple_products_fourier", "varf", "varfd",
             "tensorize", "project", "inner", "iterator_get_degree", "iterator_index", "iterator_size",
             "iterator_list_indices")
    for name in names:
        globals()[name] = cache()(debug()(log_stats()(globals()[name])))


_decorate()
