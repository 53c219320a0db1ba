"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_ref/libhermite_ref*.so, i.e. the
reference's own C++ (compiled by oracle/Makefile from /root/reference/cpp/src/hermite/*.cpp) behind
the thin C wrapper oracle/ref_capi.cpp.  Function names and argument meaning mirror
hermipy/core.py:125-348 of the reference so that parity tests read like the reference's own tests.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package must never do so.

CAUTION: the reference calls exit() on invalid input (lib.cpp:93-99, transform.cpp:94-95,166): a bad
call here kills the interpreter, exactly as it would with the original hermite_cpp module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_uint)


def available(variant=""):
    return os.path.exists(os.path.join(_HERE, "_ref", "libhermite_ref%s.so" % variant))


def lib(variant=""):
    """variant: "" (-Ofast, as the reference builds), "_o2" (strict IEEE), "_big" (bisection cap raised)."""
    if variant not in _libs:
        path = os.path.join(_HERE, "_ref", "libhermite_ref%s.so" % variant)
        L = C.CDLL(path)
        L.ref_integrate.restype = C.c_double
        for name in ("ref_index_size", "ref_index_get_degree", "ref_index_list", "ref_triangle_index",
                     "ref_cube_index"):
            getattr(L, name).restype = C.c_uint
        _libs[variant] = L
    return _libs[variant]


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _grid(nodes, weights):
    n_pts = _i([len(n) for n in nodes])
    return n_pts, _d(np.concatenate([_d(n) for n in nodes])), _d(np.concatenate([_d(w) for w in weights]))


def hermite_eval(x, degree, variant=""):
    out = np.zeros(degree + 1)
    lib(variant).ref_hermite_eval(C.c_double(x), C.c_uint(degree), _dp(out))
    return out


def discretize(function, nodes, translation, dilation, variant=""):
    nodes = [_d(n) for n in nodes]
    dim = len(nodes)
    n_pts = _i([len(n) for n in nodes])
    nd, tr, di = _d(np.concatenate(nodes)), _d(translation), _d(dilation).reshape(dim, dim)
    out = np.zeros(int(np.prod(n_pts)))
    rc = lib(variant).ref_discretize(str(function).encode(), dim, _ip(n_pts), _dp(nd), _dp(tr), _dp(di), _dp(out), out.size)
    assert rc == 0, rc
    return out


def iterator_size(dim, degree, index_set="triangle", variant=""):
    return int(lib(variant).ref_index_size(index_set.encode(), dim, degree))


def iterator_get_degree(dim, n_polys, index_set="triangle", variant=""):
    return int(lib(variant).ref_index_get_degree(index_set.encode(), dim, n_polys))


def iterator_list_indices(dim, degree, index_set="triangle", variant=""):
    L = lib(variant)
    n = int(L.ref_index_list(index_set.encode(), dim, degree, None, 0))
    out = np.zeros((n, dim), dtype=np.uint32)
    L.ref_index_list(index_set.encode(), dim, degree, out.ctypes.data_as(c_up), n)
    return out.astype(int)


def iterator_index(mult_ind, index_set="triangle", variant=""):
    m = np.ascontiguousarray(mult_ind, dtype=np.uint32)
    fn = {"triangle": lib(variant).ref_triangle_index, "cube": lib(variant).ref_cube_index}[index_set]
    return int(fn(m.ctypes.data_as(c_up), len(m)))


def transform(degree, fgrid, nodes, weights, forward, do_fourier=None, index_set="triangle", variant=""):
    dim = len(nodes)
    do_fourier = _i([0] * dim if do_fourier is None else do_fourier)
    n_pts, nd, wt = _grid(nodes, weights)
    f = _d(fgrid).ravel()
    n_out = iterator_size(dim, degree, index_set, variant) if forward else int(np.prod(n_pts))
    out = np.zeros(n_out)
    rc = lib(variant).ref_transform(C.c_uint(int(degree)), _dp(f), len(f), dim, _ip(n_pts), _dp(nd), _dp(wt),
                                    _ip(do_fourier), int(bool(forward)), index_set.encode(), _dp(out), n_out)
    assert rc == 0, rc
    return out


def integrate(fgrid, nodes, weights, do_fourier=None, variant=""):
    dim = len(nodes)
    do_fourier = _i([0] * dim if do_fourier is None else do_fourier)
    n_pts, nd, wt = _grid(nodes, weights)
    f = _d(fgrid).ravel()
    return float(lib(variant).ref_integrate(_dp(f), len(f), dim, _ip(n_pts), _dp(nd), _dp(wt), _ip(do_fourier)))


def triple_products(degree, variant=""):
    out = np.zeros((2 * degree + 1, degree + 1, degree + 1))
    lib(variant).ref_triple_products(int(degree), _dp(out))
    return out


def triple_products_fourier(degree, variant=""):
    out = np.zeros((2 * degree + 1, degree + 1, degree + 1))
    lib(variant).ref_triple_products_fourier(int(degree), _dp(out))
    return out


def varf(degree, fgrid, nodes, weights, sparse=False, do_fourier=None, index_set="triangle", variant=""):
    dim = len(nodes)
    do_fourier = _i([0] * dim if do_fourier is None else do_fourier)
    n_pts, nd, wt = _grid(nodes, weights)
    f = _d(fgrid).ravel()
    n = iterator_size(dim, degree, index_set, variant)
    L = lib(variant)
    if not sparse:
        out = np.zeros((n, n))
        rc = L.ref_varf(C.c_uint(int(degree)), _dp(f), len(f), dim, _ip(n_pts), _dp(nd), _dp(wt), _ip(do_fourier),
                        index_set.encode(), _dp(out), n)
        assert rc == 0, rc
        return out
    import scipy.sparse as ss
    cap = n * n
    rows, cols, vals = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    nnz = L.ref_varf_sp(C.c_uint(int(degree)), _dp(f), len(f), dim, _ip(n_pts), _dp(nd), _dp(wt), _ip(do_fourier),
                        index_set.encode(), _ip(rows), _ip(cols), _dp(vals), cap)
    return ss.csr_matrix((vals[:nnz], (rows[:nnz], cols[:nnz])), shape=(n, n))


def varfd(dim, degree, direction, var, do_fourier=0, index_set="triangle", variant=""):
    v = _d(var)
    out = np.zeros_like(v)
    lib(variant).ref_varfd(dim, int(degree), direction, _dp(v), v.shape[0], do_fourier, index_set.encode(), _dp(out))
    return out


def _dirs_args(dirs_list):
    flat = _i(np.concatenate([np.asarray(list(d), dtype=np.int32) for d in dirs_list]))
    lens = _i([len(d) for d in dirs_list])
    return flat, lens


def tensorize(inp, dim=None, direction=None, index_set="triangle", variant=""):
    """Dense subset of hermipy.core.tensorize (core.py:201-253): dict{dirs->array}, list, or (array, dim, dir)."""
    L = lib(variant)
    if isinstance(inp, dict):
        dirs_list = [sorted(k) for k in inp.keys()]
        arrays = [_d(v) for v in inp.values()]
    elif isinstance(inp, list):
        dirs_list = [[i] for i in range(len(inp))]
        arrays = [_d(v) for v in inp]
    else:
        a = _d(inp)
        if a.ndim == 1:
            n = iterator_size(dim, len(a) - 1, index_set, variant)
            out = np.zeros(n)
            rc = L.ref_tensorize_vec_id(_dp(a), len(a), dim, direction, index_set.encode(), _dp(out), n)
            assert rc == 0, rc
            return out
        n = iterator_size(dim, a.shape[0] - 1, index_set, variant)
        out = np.zeros((n, n))
        rc = L.ref_tensorize_mat_id(_dp(a), a.shape[0], dim, direction, index_set.encode(), _dp(out), n)
        assert rc == 0, rc
        return out
    tot_dim = sum(len(d) for d in dirs_list)
    dflat, dlens = _dirs_args(dirs_list)
    degree = iterator_get_degree(len(dirs_list[0]), arrays[0].shape[0], index_set, variant)
    n = iterator_size(tot_dim, degree, index_set, variant)
    if arrays[0].ndim == 1:
        flat = _d(np.concatenate(arrays))
        lens = _i([len(a) for a in arrays])
        out = np.zeros(n)
        rc = L.ref_tensorize_vecs_dirs(len(arrays), _dp(flat), _ip(lens), _ip(dflat), _ip(dlens),
                                       index_set.encode(), _dp(out), n)
    else:
        flat = _d(np.concatenate([a.ravel() for a in arrays]))
        sizes = _i([a.shape[0] for a in arrays])
        out = np.zeros((n, n))
        rc = L.ref_tensorize_mats_dirs(len(arrays), _dp(flat), _ip(sizes), _ip(dflat), _ip(dlens),
                                       index_set.encode(), _dp(out), n)
    assert rc == 0, rc
    return out


def project(inp, dim, directions, index_set="triangle", variant=""):
    if isinstance(directions, int):
        directions = [directions]
    a = _d(inp)
    dirs = _i(directions)
    degree = iterator_get_degree(dim, a.shape[0], index_set, variant)
    n = iterator_size(len(dirs), degree, index_set, variant)
    L = lib(variant)
    if a.ndim == 1:
        out = np.zeros(n)
        rc = L.ref_project_vec(_dp(a), len(a), dim, _ip(dirs), len(dirs), index_set.encode(), _dp(out), n)
    else:
        out = np.zeros((n, n))
        rc = L.ref_project_mat(_dp(a), a.shape[0], dim, _ip(dirs), len(dirs), index_set.encode(), _dp(out), n)
    assert rc == 0, rc
    return out


def inner(s1, s2, d1, d2, index_set="triangle", variant=""):
    s1, s2, d1, d2 = _d(s1), _d(s2), _i(d1), _i(d2)
    dim1, dim2 = len(d1), len(d2)
    degree = iterator_get_degree(dim1, len(s1), index_set, variant)
    n_res = len(set(d1.tolist()) ^ set(d2.tolist()))
    n = iterator_size(n_res, degree, index_set, variant) if n_res > 0 else 1
    out = np.zeros(n)
    rc = lib(variant).ref_inner(_dp(s1), len(s1), _dp(s2), len(s2), _ip(d1), dim1, _ip(d2), dim2,
                                index_set.encode(), _dp(out), n)
    assert rc == 0, rc
    return out
