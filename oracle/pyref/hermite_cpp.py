"""TEST INFRASTRUCTURE ONLY -- a stand-in for the reference's Boost.Python module `hermite_cpp`
(cpp/src/hermite/python/module.cpp:45-168), implemented over oracle/_ref (the reference's own C++
numerical code, compiled by oracle/Makefile) through oracle/ref.py.

With this directory and /root/reference on sys.path (and a matplotlib stub), the reference's UNMODIFIED
Python package `hermipy` imports and runs in this container, which is how oracle/make_golden_api.py
produces golden vectors for the object layer (Quad / Series / Varf).  Only the part of the module
surface that hermipy/core.py touches is provided: container classes that behave like the
vector_indexing_suite wrappers (list semantics), dense/sparse matrix handles, and the functions.
"""
import os
import sys

import numpy as np
import scipy.sparse as ss

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref  # noqa: E402

VARIANT = ""          # the -Ofast build, as the reference's CMakeLists.txt:8


class double_vec(list):
    pass


class int_vec(list):
    pass


class double_mat(list):
    pass


class int_mat(list):
    pass


class double_cube(list):
    pass


class boost_mat:
    def __init__(self, a):
        self.a = np.ascontiguousarray(a, dtype=float)

    def size1(self):
        return self.a.shape[0]

    def size2(self):
        return self.a.shape[1]


class sparse_matrix:
    def __init__(self, m):
        self.m = ss.csr_matrix(m)

    def size1(self):
        return self.m.shape[0]

    def size2(self):
        return self.m.shape[1]


# ---- converters -------------------------------------------------------------------------------------
def to_numpy(x):
    if isinstance(x, boost_mat):
        return x.a.copy()
    return np.array(x, dtype=float)


def to_boost_mat(a):
    return boost_mat(a)


def to_spmat(data, indices, indptr, size1, size2):
    return sparse_matrix(ss.csr_matrix((np.array(data, float), np.array(indices, np.int64), np.array(indptr, np.int64)),
                                       shape=(size1, size2)))


def row_col_val(sp):
    coo = sp.m.tocoo()
    return [list(coo.row.astype(float)), list(coo.col.astype(float)), list(coo.data)]


def full(sp):
    return boost_mat(sp.m.toarray())


def _arr(x):
    return np.array(x, dtype=float)


# ---- functions --------------------------------------------------------------------------------------
def discretize(body, nodes, translation, dilation):
    return double_vec(ref.discretize(body, [_arr(n) for n in nodes], _arr(translation), _arr(dilation), variant=VARIANT))


def integrate(fgrid, nodes, weights, do_fourier):
    return ref.integrate(_arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights], list(do_fourier),
                         variant=VARIANT)


def transform(degree, fgrid, nodes, weights, do_fourier, forward, index_set):
    return double_vec(ref.transform(degree, _arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights], forward,
                                    do_fourier=list(do_fourier), index_set=index_set, variant=VARIANT))


def triple_products(degree):
    return ref.triple_products(degree, variant=VARIANT)


def triple_products_fourier(degree):
    return ref.triple_products_fourier(degree, variant=VARIANT)


def varf(degree, fgrid, nodes, weights, do_fourier, index_set):
    return boost_mat(ref.varf(degree, _arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights],
                              do_fourier=list(do_fourier), index_set=index_set, variant=VARIANT))


def varf_sp(degree, fgrid, nodes, weights, do_fourier, index_set):
    return sparse_matrix(ref.varf(degree, _arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights], sparse=True,
                                  do_fourier=list(do_fourier), index_set=index_set, variant=VARIANT))


def varfd(dim, degree, direction, var, do_fourier, index_set):
    sparse = isinstance(var, sparse_matrix)
    a = var.m.toarray() if sparse else var.a
    out = ref.varfd(dim, degree, direction, a, do_fourier=do_fourier, index_set=index_set, variant=VARIANT)
    return sparse_matrix(out) if sparse else boost_mat(out)


def _tensorize(args, sparse):
    index_set = args[-1]
    if len(args) == 3:                       # (arrays, dirs, index_set)
        arrays, dirs = args[0], args[1]
        inp = {frozenset(d): _arr(a) for a, d in zip(arrays, dirs)}
        out = ref.tensorize(inp, index_set=index_set, variant=VARIANT)
    elif len(args) == 2:                     # (arrays, index_set): one factor per axis
        out = ref.tensorize([_arr(a) for a in args[0]], index_set=index_set, variant=VARIANT)
    else:                                    # (array, dim, direction, index_set)
        out = ref.tensorize(_arr(args[0]), dim=args[1], direction=args[2], index_set=index_set, variant=VARIANT)
    if out.ndim == 1:
        return double_vec(out)
    return sparse_matrix(out) if sparse else double_mat([list(r) for r in out])


def tensorize(*args):
    return _tensorize(args, False)


def tensorize_sp(*args):
    return _tensorize(args, True)


def project(inp, dim, directions, index_set):
    sparse = isinstance(inp, sparse_matrix)
    a = inp.m.toarray() if sparse else (inp.a if isinstance(inp, boost_mat) else _arr(inp))
    out = ref.project(a, dim, list(directions), index_set=index_set, variant=VARIANT)
    if out.ndim == 1:
        return double_vec(out)
    return sparse_matrix(out) if sparse else double_mat([list(r) for r in out])


def inner(s1, s2, d1, d2, index_set):
    return double_vec(ref.inner(_arr(s1), _arr(s2), list(d1), list(d2), index_set=index_set, variant=VARIANT))


def triangle_index(m):
    return ref.iterator_index(list(m), "triangle", variant=VARIANT)


def cube_index(m):
    return ref.iterator_index(list(m), "cube", variant=VARIANT)


def _per_set(name):
    g = globals()
    g[name + "_get_degree"] = lambda dim, n_polys: ref.iterator_get_degree(dim, n_polys, name, variant=VARIANT)
    g[name + "_size"] = lambda dim, degree: ref.iterator_size(dim, degree, name, variant=VARIANT)
    g[name + "_list_indices"] = lambda dim, degree: ref.iterator_list_indices(dim, degree, name, variant=VARIANT)


for _n in ("triangle", "cross", "cross_nc", "cube", "rectangle"):
    _per_set(_n)
