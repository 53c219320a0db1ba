"""TEST INFRASTRUCTURE.  Swaps the compute functions of hermipy_b200.core for the reference's own C++
(oracle/_ref through oracle/ref.py) so that the host-side object layer (Quad / Series / Varf / Function /
tensorize_at / cache) can be exercised on a machine without a GPU.  The index-set functions are NOT
swapped: they are host code of libhermite_b200.so and run anywhere."""
import contextlib

import numpy as np
import scipy.sparse as ss

from hermipy_b200 import core
from oracle import ref

_VARIANT = "_big"     # same sources, DEGREE_BISSECT_MAX raised (oracle/Makefile)


def _dense(a):
    return np.asarray(a.todense()) if ss.issparse(a) else np.asarray(a, dtype=float)


def _varf(degree, fgrid, nodes, weights, sparse=False, do_fourier=None, index_set="triangle", mode="reference"):
    return ref.varf(degree, fgrid, nodes, weights, sparse=sparse, do_fourier=do_fourier, index_set=index_set,
                    variant=_VARIANT)


def _varfd(dim, degree, direction, var, do_fourier=0, index_set="triangle"):
    out = ref.varfd(dim, degree, direction, _dense(var), do_fourier=do_fourier, index_set=index_set, variant=_VARIANT)
    return ss.csr_matrix(out) if ss.issparse(var) else out


def _tensorize(inp, dim=None, direction=None, sparse=False, index_set="triangle"):
    if isinstance(inp, dict):
        first = next(iter(inp.values()))
        if isinstance(first, (float, int)):
            return np.prod([inp[k] for k in inp])
        inp = {k: _dense(v) for k, v in inp.items()}
    elif isinstance(inp, list):
        if isinstance(inp[0], (float, int)):
            return np.prod(inp)
        inp = [_dense(v) for v in inp]
    else:
        inp = _dense(inp)
    out = ref.tensorize(inp, dim=dim, direction=direction, index_set=index_set, variant=_VARIANT)
    return ss.csr_matrix(out) if (sparse and out.ndim == 2) else out


def _project(inp, dim, directions, index_set="triangle"):
    out = ref.project(_dense(inp), dim, directions, index_set=index_set, variant=_VARIANT)
    return ss.csr_matrix(out) if ss.issparse(inp) else out


def _with_variant(fn):
    def call(*args, **kwargs):
        return fn(*args, variant=_VARIANT, **kwargs)
    call.__name__ = fn.__name__
    return call


def _grid_points(nodes, translation, dilation):
    mesh = np.meshgrid(*[np.asarray(n, float) for n in nodes], indexing="ij")
    pts = np.stack([m.ravel() for m in mesh], axis=1)
    return np.asarray(translation, float).ravel() + pts @ np.asarray(dilation, float).reshape(len(nodes), len(nodes)).T


def _expr(body, nodes, translation, dilation):
    return ref.discretize(body, nodes, translation, dilation, variant=_VARIANT)


def _transform_weighted(degree, function, factor, nodes, weights, translation, dilation, do_fourier=None,
                        index_set="triangle"):
    """host restatement of hermipy/quad.py:287-306 on the reference's own transform (test infrastructure)"""
    n_grid = int(np.prod([len(n) for n in nodes]))
    if isinstance(function, str):
        f, single = _expr(function, nodes, translation, dilation)[None], True
    else:
        f = np.asarray(function, float)
        single = f.size == n_grid and (f.ndim != 2 or f.shape[0] != 1 or n_grid == 1)
        f = f.reshape(-1, n_grid)
    if factor is not None:
        fac = _expr(factor, nodes, translation, dilation)
        f = f / (1e-300 * (fac == 0) + fac)
    out = np.stack([ref.transform(degree, row, nodes, weights, True, do_fourier=do_fourier, index_set=index_set,
                                  variant=_VARIANT) for row in f])
    return out[0] if single else out


def _eval_weighted(degree, coeffs, factor, eval_nodes, nodes, weights, translation, dilation, do_fourier=None,
                   index_set="triangle"):
    values = ref.transform(degree, coeffs, eval_nodes, weights, False, do_fourier=do_fourier, index_set=index_set,
                           variant=_VARIANT)
    return values if factor is None else values * _expr(factor, nodes, translation, dilation)


def _varf_function(degree, function, nodes, weights, translation, dilation, sparse=False, do_fourier=None,
                   index_set="triangle", mode="reference"):
    return _varf(degree, _expr(function, nodes, translation, dilation), nodes, weights, sparse=sparse,
                 do_fourier=do_fourier, index_set=index_set)


PATCHES = {
    "transform_weighted": _transform_weighted,
    "eval_weighted": _eval_weighted,
    "varf_function": _varf_function,
    "discretize": _with_variant(ref.discretize),
    "integrate": _with_variant(ref.integrate),
    "transform": _with_variant(ref.transform),
    "triple_products": _with_variant(ref.triple_products),
    "varf": _varf,
    "varfd": _varfd,
    "tensorize": _tensorize,
    "project": _project,
    "inner": _with_variant(ref.inner),
}


@contextlib.contextmanager
def oracle_core():
    saved = {name: getattr(core, name) for name in PATCHES}
    try:
        for name, fn in PATCHES.items():
            setattr(core, name, fn)
        yield
    finally:
        for name, fn in saved.items():
            setattr(core, name, fn)
