"""Quadrature object: the user-facing entry point of the hot path (reference: hermipy/quad.py:32-425).

`Quad.transform`, `Quad.eval`, `Quad.varf`, `Quad.varfd`, `Quad.discretize` and `Quad.integrate` keep the
reference's signatures and semantics; the arithmetic behind them runs in libhermite_b200.so:

    discretize -> NVRTC-compiled expression evaluated on the GPU grid        (core.discretize)
    transform  -> one FP64 tensor-core GEMM per axis (sum-factorisation)     (core.transform)
    varf       -> thresholded triple-product assembly / two-stage DMMA GEMM  (core.varf)
    varfd, tensorize, project -> index-map gather kernels

Host-side here: the pre/post scaling by the weight factor (quad.py:297-298, :328), the affine node map of
`eval` (:315-322) and the `tensorize_at` decorator (:177-240) which splits separable functions into
lower-dimensional problems and recombines them with core.tensorize / core.inner.
"""
import functools

import numpy as np
import numpy.linalg as la
import sympy as sym

from . import core
from . import lib
from . import stats
from .function import Function
from .operator import Operator
from .position import Position
from .series import Series
from .settings import settings
from .varf import Varf

very_small = 1e-10


def _stack(arrays):
    """(dim, n) array when every axis has the same number of points, else a list of 1-D arrays."""
    arrays = [np.asarray(a, float) for a in arrays]
    if len({len(a) for a in arrays}) == 1:
        return np.asarray(arrays, float)
    return arrays


def tensorize_at(arg_num):
    """Decorator for Quad methods whose argument `arg_num` is a function of the space variables.  When the
    function is a sum of products of factors living on disjoint groups of directions (and the quadrature's
    own weight factor splits the same way), the method is applied to every group on the projected
    quadrature and the partial results are tensorised, instead of working on the full grid."""
    def decorate(method):
        @functools.wraps(method)
        def wrapper(*args, **kwargs):
            requested = kwargs.get('tensorize')
            enabled = settings['tensorize'] if requested is None else requested
            quad, function = args[0], args[arg_num]
            if isinstance(function, Series):      # integrate(series): evaluated on the full grid
                enabled = False
            if kwargs.get('resident'):            # device-resident results are assembled on the full grid
                enabled = False
            if not enabled or not quad.position.is_diag or isinstance(function, np.ndarray):
                return method(*args, **kwargs)
            if not isinstance(function, Function):
                function = Function(function, dirs=quad.position.dirs)
            weight_groups = quad.factor.split(legacy=False)
            if len(weight_groups) != 1:
                raise ValueError("Tensorization not possible!")
            pass_on = {'sparse': kwargs['sparse']} if 'sparse' in kwargs else {}
            total = None
            for term in function.split(legacy=False):
                multiplier = term.pop(frozenset())
                # coarsest grouping of directions under which both the term and the weight are products
                blocks = lib.finest_common(set(term), set(weight_groups[0]))
                partial = {}
                for block in blocks:
                    piece = sym.Integer(1)
                    for dirs, part in term.items():
                        if dirs <= block:
                            piece = piece * part.sym
                    sub_args = list(args)
                    sub_args[0] = quad.project(sorted(block))
                    sub_args[arg_num] = Function(piece, dirs=sorted(block))
                    partial[block] = method(*sub_args, **kwargs)
                if settings['debug']:
                    print("Tensorizing results")
                kind = type(next(iter(partial.values())))
                if kind is Varf or kind is Series:
                    combined = kind.tensorize(list(partial.values()), **pass_on)
                else:
                    combined = core.tensorize(partial, **pass_on)
                combined = combined * float(multiplier.sym)
                total = combined if total is None else total + combined
            return total
        return wrapper
    return decorate


class Quad:

    @staticmethod
    def tensorize(args):
        if not len(args) > 0 or any(not isinstance(a, Quad) for a in args):
            raise ValueError("Invalid argument(s)!")
        position = Position.tensorize([a.position for a in args])
        nodes, weights = [0] * len(position.dirs), [0] * len(position.dirs)
        factor = sym.Integer(1)
        for a in args:
            factor = factor * a.factor.sym
            for i, d in enumerate(a.position.dirs):
                nodes[position.dirs.index(d)] = a.nodes[i]
                weights[position.dirs.index(d)] = a.weights[i]
        return Quad(nodes, weights, position, Function(factor, dirs=position.dirs))

    def __init__(self, nodes, weights, position=None, factor=1, mean=None, cov=None, dirs=None, types=None):
        """nodes / weights: one 1-D array per direction (weights already normalised for the Gaussian of the
        basis).  `factor` is the weight function between the solution and its series: transform() divides by
        it, eval() multiplies by it."""
        self.nodes = _stack(nodes)
        self.weights = _stack(weights)
        dim = len(self.nodes)
        self.position = position if position is not None else \
            Position(dim=dim, mean=mean, cov=cov, dirs=dirs, types=types)
        self.factor = Function(factor, dirs=self.position.dirs)
        self.factor.sym = self.factor.sym.expand()
        self.do_fourier = [int(t == "fourier") for t in self.position.types]

    # ---- constructors ---------------------------------------------------------------------------
    @classmethod
    def gauss_hermite(cls, n_points, dim=None, dirs=None, **kwargs):
        if dirs is not None:
            dim = len(dirs)
        if dim is not None:
            n_points = np.full(dim, n_points)
        elif isinstance(n_points, (int, np.integer)):
            n_points = [n_points]
        nodes, weights = lib.hermegauss_nd(n_points)
        return cls(nodes, weights, dirs=dirs, **kwargs)

    @classmethod
    def fourier(cls, n_points, bounds=None, position=None, factor=1, dirs=None):
        if position is not None:
            dim = position.dim
        if dirs is not None:
            dim = len(dirs)
        elif bounds is not None:
            dim = len(bounds)
        elif not isinstance(n_points, (int, np.integer)):
            dim = len(n_points)
        elif position is None:
            dim = 1
        if position is None:
            bounds = [(-np.pi, np.pi)] * dim if bounds is None else bounds
            dirs = list(range(dim)) if dirs is None else dirs
            if dim != len(bounds) or dim != len(dirs):
                raise ValueError("Incompatible arguments!")
            # (mean, cov) encode the interval of a Fourier direction: midpoint and (length / 2 pi)^2
            mean = [left / 2 + right / 2 for left, right in bounds]
            cov = np.diag([((right - left) / (2 * np.pi)) ** 2 for left, right in bounds])
            position = Position(dim=dim, dirs=dirs, mean=mean, cov=cov, types=["fourier"] * dim)
        if isinstance(n_points, (int, np.integer)):
            n_points = [n_points] * dim
        assert dim == len(n_points)
        nodes = [-np.pi + 2 * np.pi * np.arange(n) / n for n in n_points]
        weights = [2 * np.pi / n * np.ones(n) for n in n_points]
        return cls(nodes, weights, position=position, factor=factor)

    @classmethod
    def newton_cotes(cls, n_points, extrema, **kwargs):
        """Composite trapezoidal rule on [-extremum, extremum] per direction, against the Gaussian weight."""
        nodes, weights = [], []
        for n, extremum in zip(n_points, extrema):
            x = np.linspace(-extremum, extremum, n)
            rule = np.ones(n)
            rule[0] = rule[-1] = .5
            gaussian = np.exp(-x ** 2 / 2.) / np.sqrt(2 * np.pi)
            nodes.append(x)
            weights.append(rule * gaussian * (2 * extremum / (n - 1)))
        return cls(nodes, weights, **kwargs)

    # ---- algebra --------------------------------------------------------------------------------
    def __mul__(self, other):
        return Quad.tensorize([self, other])

    def __eq__(self, other):
        if not isinstance(other, Quad):
            raise ValueError("Invalid argument!")
        close = lambda a, b: len(a) == len(b) and all(
            len(u) == len(v) and la.norm(np.asarray(u) - np.asarray(v)) < very_small for u, v in zip(a, b))
        return self.position == other.position and self.factor == other.factor \
            and close(self.nodes, other.nodes) and close(self.weights, other.weights)

    def __str__(self):
        return "Quadrature in dimension " + str(self.position.dim) \
               + "\n- Associated factor: " + str(self.factor) \
               + "\n- Types: " + str(self.position.types)

    __repr__ = __str__

    tensorize_at = staticmethod(tensorize_at)

    # ---- grid functions -------------------------------------------------------------------------
    def mapped_nodes(self):
        """Physical coordinates of every grid point, shape (n_points_total, dim)."""
        columns = [self.discretize(Function.x_sub[d]) for d in self.position.dirs]
        return np.asarray(np.vstack(columns)).T

    def discretize(self, f):
        """Values of f at the physical grid points mean + factor @ node, flat, last axis fastest."""
        if not isinstance(f, Function):
            f = Function(f, dirs=self.position.dirs)
        return core.discretize(f.ccode(), self.nodes, self.position.mean, self.position.factor)

    @tensorize_at(1)
    def integrate(self, function, flat=False, tensorize=None):
        """Integral against the Gaussian weight of the quadrature, or against Lebesgue measure if `flat`."""
        if isinstance(function, Series):
            function = self.eval(function)
        if not isinstance(function, np.ndarray):
            function = self.discretize(function)
        if flat:
            w_grid = self.discretize(self.position.weight())
            w_grid = 1e-300 * (w_grid == 0) + w_grid       # outer nodes where the Gaussian underflows
            function = function / w_grid
        return core.integrate(function, self.nodes, self.weights, do_fourier=self.do_fourier)

    def norm(self, function, n=2, flat=False):
        if isinstance(function, Series):
            function = self.eval(function)
        if n == 2:
            return np.sqrt(self.integrate(function ** 2, flat=flat))
        elif n == 1:
            return self.integrate(abs(function), flat=flat)

    # ---- the hot path ---------------------------------------------------------------------------
    @tensorize_at(1)
    @stats.debug()
    @stats.log_stats()
    def transform(self, function, degree, index_set="triangle", significant=0, tensorize=None):
        """Forward transform: coefficients of function / factor in the basis of this quadrature."""
        # one device pass: the function (unless given as grid values) and the factor are evaluated on the GPU grid,
        # divided there and transformed; only the coefficients come back (hermipy/quad.py:287-306 does these steps
        # with NumPy on the host)
        if not isinstance(function, np.ndarray):
            function = Function(function, dirs=self.position.dirs).ccode()
        else:
            function = np.asarray(function, float).ravel()
        coeffs = core.transform_weighted(degree, function, self.factor.ccode(), self.nodes, self.weights,
                                         self.position.mean, self.position.factor, do_fourier=self.do_fourier,
                                         index_set=index_set)
        return Series(coeffs, self.position, factor=self.factor, index_set=index_set, significant=significant)

    def transform_batch(self, functions, degree, index_set="triangle"):
        """Forward transform of many grid functions (array of shape (batch, n_points_total)) in one device
        call; returns a list of Series.  No reference equivalent (the reference loops in Python)."""
        functions = np.asarray(functions, float)
        functions = functions.reshape(functions.shape[0], -1)
        coeffs = core.transform_weighted(degree, functions, self.factor.ccode(), self.nodes, self.weights,
                                         self.position.mean, self.position.factor, do_fourier=self.do_fourier,
                                         index_set=index_set)
        coeffs = np.atleast_2d(coeffs)
        return [Series(c, self.position, factor=self.factor, index_set=index_set) for c in coeffs]

    def eval(self, series):
        """Inverse transform: values of the series (times its factor) at the grid points of THIS
        quadrature, which may be centred/scaled differently from the series' own position."""
        if isinstance(series, np.ndarray):
            series = Series(series, self.position, factor=self.factor)
        inv = la.inv(series.position.factor)
        translation = inv.dot(self.position.mean - series.position.mean)
        scaling = inv * self.position.factor          # entry-wise: only diagonal maps are supported
        if la.norm(scaling - np.diag(np.diag(scaling)), 2) > 1e-8:
            raise ValueError("Incompatible covariance matrices")
        mapped_nodes = [np.asarray(n) * scaling[i][i] + translation[i] for i, n in enumerate(self.nodes)]
        factor = series.factor if isinstance(series.factor, Function) else Function(series.factor, dirs=self.position.dirs)
        return core.eval_weighted(series.degree, series.coeffs, factor.ccode(), mapped_nodes, self.nodes, self.weights,
                                  self.position.mean, self.position.factor, do_fourier=self.do_fourier,
                                  index_set=series.index_set)

    @tensorize_at(1)
    @stats.debug()
    @stats.log_stats()
    def varf(self, f_grid, degree, sparse=None, index_set="triangle", tensorize=None, resident=False):
        """Matrix of the multiplication operator u -> f u in the basis of this quadrature.  `resident=True`
        leaves the (dense) matrix in HBM and returns a `ResidentVarf` (hermipy_b200/resident.py)."""
        sparse = settings['sparse'] if sparse is None else sparse
        if not isinstance(f_grid, np.ndarray) and not resident:
            # the expression is evaluated on the device grid and assembled there: no grid values cross PCIe
            f = f_grid if isinstance(f_grid, Function) else Function(f_grid, dirs=self.position.dirs)
            matrix = core.varf_function(degree, f.ccode(), self.nodes, self.weights, self.position.mean,
                                        self.position.factor, sparse=sparse, do_fourier=self.do_fourier,
                                        index_set=index_set)
            return Varf(matrix, self.position, factor=self.factor, index_set=index_set)
        if not isinstance(f_grid, np.ndarray):
            f_grid = self.discretize(f_grid)
        if resident:
            from .resident import ResidentVarf
            plan = self._resident_plan(degree, index_set)
            block = plan.varf(plan.pack_grid(np.asarray(f_grid, float))[0], mode="reference")
            return ResidentVarf(block, self.position, factor=self.factor, index_set=index_set)
        matrix = core.varf(degree, f_grid, self.nodes, self.weights, sparse=sparse,
                           do_fourier=self.do_fourier, index_set=index_set)
        return Varf(matrix, self.position, factor=self.factor, index_set=index_set)

    def _resident_plan(self, degree, index_set):
        """Device plan (basis / triple-product tables) of this quadrature for one (degree, index set)."""
        from .plan import Plan
        plans = self.__dict__.setdefault('_plans', {})
        key = (int(degree), index_set)
        if key not in plans:
            plans[key] = Plan(degree, self.nodes, self.weights, do_fourier=self.do_fourier, index_set=index_set)
        return plans[key]

    @tensorize_at(1)
    @stats.debug()
    @stats.log_stats()
    def varfd(self, function, degree, directions, sparse=None, index_set="triangle", tensorize=None,
              resident=False):
        """Matrix of u -> function * d^k u / (dx_{d1} ... dx_{dk}), `directions` = [d1, ..., dk]."""
        directions = [d for d in directions if d in self.position.dirs]
        sparse = settings['sparse'] if sparse is None else sparse
        eigval, _ = la.eig(self.position.cov)
        if resident:
            from .plan import varfd_device
            from .resident import ResidentMatrix, ResidentVarf
            block = self.varf(function, degree, index_set=index_set, tensorize=False, resident=True).matrix.block
            for d in directions:
                rel = self.position.dirs.index(d)
                block = varfd_device(self.position.dim, degree, rel, block, do_fourier=self.do_fourier[rel],
                                     index_set=index_set)
                block /= float(np.sqrt(eigval[rel]))
            return ResidentVarf(ResidentMatrix(block), self.position, factor=self.factor, index_set=index_set)
        matrix = self.varf(function, degree, sparse=sparse, index_set=index_set, tensorize=False).matrix
        for d in directions:
            rel = self.position.dirs.index(d)
            matrix = core.varfd(self.position.dim, degree, rel, matrix, do_fourier=self.do_fourier[rel],
                                index_set=index_set)
            matrix = matrix / np.sqrt(eigval[rel])
        return Varf(matrix, self.position, factor=self.factor, index_set=index_set)

    @stats.debug()
    @stats.log_stats()
    def discretize_op(self, op, degree, sparse=None, index_set="triangle", resident=False):
        """Galerkin matrix of a linear differential operator (conjugated by the weight factor)."""
        if not isinstance(op, Operator):
            op = Operator(op, dirs=self.position.dirs)
        if self.factor != Function(1, dirs=self.position.dirs):
            op = op.map(self.factor)
        if op.dirs != self.position.dirs:
            raise ValueError("Invalid argument: directions don't match")
        sparse = settings['sparse'] if sparse is None else sparse
        extra = {'resident': True} if resident else {}
        terms = op.split()
        if terms == {}:
            return self.varf(0, degree, sparse=sparse, index_set=index_set, **extra)
        total = 0
        for orders, coefficient in terms.items():
            derivatives = [d for i, d in enumerate(self.position.dirs) for _ in range(orders[i])]
            total = total + self.varfd(coefficient, degree, derivatives, sparse=sparse, index_set=index_set, **extra)
        return total

    # ---- bookkeeping ----------------------------------------------------------------------------
    def project(self, directions):
        """Quadrature restricted to a sorted list of directions (weight factor must split accordingly)."""
        factor = self.factor.project(directions)
        position = self.position.project(directions)
        rel = [self.position.dirs.index(d) for d in position.dirs]
        return Quad([self.nodes[r] for r in rel], [self.weights[r] for r in rel], position=position, factor=factor)

    def series(self, coeffs, index_set="triangle"):
        return Series(coeffs, self.position, factor=self.factor, index_set=index_set)

    def plot(self, arg, *args, **kwargs):
        from ._plotting import plot_quad
        return plot_quad(self, arg, *args, **kwargs)

    def plot_hf(self, multi_index, ax=None, bounds=True, **kwargs):
        from ._plotting import plot_hermite_function
        return plot_hermite_function(self, multi_index, ax=ax, bounds=bounds, **kwargs)
