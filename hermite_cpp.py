"""`hermite_cpp` -- the module name of the reference's Boost.Python bridge (cpp/src/hermite/python/module.cpp:45-168),
implemented over libhermite_b200.so.  With this file (the repository root) on sys.path BEFORE any other
`hermite_cpp`, the reference's own, unmodified `hermipy/core.py` -- hence its whole Python package -- runs on the CUDA
library: every function below forwards to `hermipy_b200.core` (the ctypes mirror of the C ABI).  The container
classes behave like the vector_indexing_suite wrappers the reference builds its arguments with (list semantics,
`extend` / `append`); matrices cross as NumPy arrays / SciPy CSR inside thin handles.  Converters of
module.cpp:123-135 (`to_numpy`, `to_mat`, `to_bmat`, `to_boost_mat`, `full`, `to_spmat`, `row_col_val`) are host-side
format changes, as in the reference.
"""
import numpy as np
import scipy.sparse as ss

from hermipy_b200 import core as _core


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
    """dense row-major matrix handle (boost::numeric::ublas::matrix<double> in the reference)"""

    def __init__(self, a):
        self.a = np.ascontiguousarray(a, dtype=np.float64)

    def size1(self):
        return self.a.shape[0]

    def size2(self):
        return self.a.shape[1]


class sparse_matrix:
    """CSR handle (ublas compressed_matrix in the reference)"""

    def __init__(self, m):
        self.m = m if isinstance(m, ss.csr_matrix) else ss.csr_matrix(m)

    def size1(self):
        return self.m.shape[0]

    def size2(self):
        return self.m.shape[1]


# ---- converters (module.cpp:123-135) ------------------------------------------------------------------
def to_numpy(x):
    if isinstance(x, boost_mat):
        return x.a.copy()
    return np.array(x, dtype=np.float64)


def to_mat(a):
    out = double_mat()
    for row in np.asarray(a, dtype=np.float64):
        v = double_vec()
        v.extend(row.tolist())
        out.append(v)
    return out


def to_bmat(a):
    return boost_mat(a)


def to_boost_mat(a):
    return boost_mat(a)


def to_spmat(data, indices, indptr, size1, size2):
    return sparse_matrix(ss.csr_matrix((np.array(data, dtype=np.float64), np.array(indices, dtype=np.int64),
                                        np.array(indptr, dtype=np.int64)), shape=(size1, size2)))


def row_col_val(sp):
    coo = sp.m.tocoo()
    return [list(coo.row.astype(float)), list(coo.col.astype(float)), list(coo.data)]


def full(sp):
    return boost_mat(sp.m.toarray())


def _arr(x):
    return np.array(x, dtype=np.float64)


def _dense(x):
    if isinstance(x, boost_mat):
        return x.a
    if isinstance(x, sparse_matrix):
        return x.m
    return _arr(x)


# ---- functions (module.cpp:88-131): same argument order as the Boost.Python signatures ---------------------
def discretize(body, nodes, translation, dilation):
    return double_vec(_core.discretize(body, [_arr(n) for n in nodes], _arr(translation), _arr(dilation)))


def integrate(fgrid, nodes, weights, do_fourier):
    return _core.integrate(_arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights], do_fourier=list(do_fourier))


def transform(degree, fgrid, nodes, weights, do_fourier, forward, index_set):
    return double_vec(_core.transform(degree, _arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights], forward,
                                      do_fourier=list(do_fourier), index_set=index_set))


def triple_products(degree):
    return _core.triple_products(degree)


def triple_products_fourier(degree):
    return _core.triple_products_fourier(degree)


def varf(degree, fgrid, nodes, weights, do_fourier, index_set):
    return boost_mat(_core.varf(degree, _arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights],
                                do_fourier=list(do_fourier), index_set=index_set))


def varf_sp(degree, fgrid, nodes, weights, do_fourier, index_set):
    return sparse_matrix(_core.varf(degree, _arr(fgrid), [_arr(n) for n in nodes], [_arr(w) for w in weights], sparse=True,
                                    do_fourier=list(do_fourier), index_set=index_set))


def varfd(dim, degree, direction, var, do_fourier, index_set):
    out = _core.varfd(dim, degree, direction, _dense(var), do_fourier=do_fourier, index_set=index_set)
    return sparse_matrix(out) if isinstance(var, sparse_matrix) else boost_mat(out)


def _tensorize(args, sparse):
    index_set = args[-1]
    if len(args) == 3:                       # (arrays, dirs, index_set)
        inp = {frozenset(d): _dense(a) for a, d in zip(args[0], args[1])}
        out = _core.tensorize(inp, sparse=sparse, index_set=index_set)
    elif len(args) == 2:                     # (arrays, index_set): one factor per axis
        out = _core.tensorize([_dense(a) for a in args[0]], sparse=sparse, index_set=index_set)
    else:                                    # (array, dim, direction, index_set)
        out = _core.tensorize(_dense(args[0]), dim=args[1], direction=args[2], sparse=sparse, index_set=index_set)
    if isinstance(out, ss.csr_matrix):
        return sparse_matrix(out)
    if out.ndim == 1:
        return double_vec(out)
    return double_mat([list(r) for r in out])


def tensorize(*args):
    return _tensorize(args, False)


def tensorize_sp(*args):
    return _tensorize(args, True)


def project(inp, dim, directions, index_set):
    out = _core.project(_dense(inp), dim, list(directions), index_set=index_set)
    if isinstance(out, ss.csr_matrix):
        return sparse_matrix(out)
    if out.ndim == 1:
        return double_vec(out)
    return double_mat([list(r) for r in out])


def inner(s1, s2, d1, d2, index_set):
    return double_vec(_core.inner(_arr(s1), _arr(s2), list(d1), list(d2), index_set=index_set))


def triangle_index(m):
    return _core.iterator_index(list(m), "triangle")


def cube_index(m):
    return _core.iterator_index(list(m), "cube")


def _per_set(name):
    g = globals()
    g[name + "_get_degree"] = lambda dim, n_polys: _core.iterator_get_degree(dim, n_polys, name)
    g[name + "_size"] = lambda dim, degree: _core.iterator_size(dim, degree, name)
    g[name + "_list_indices"] = lambda dim, degree: _core.iterator_list_indices(dim, degree, name)


for _n in ("triangle", "cross", "cross_nc", "cube", "rectangle"):
    _per_set(_n)
