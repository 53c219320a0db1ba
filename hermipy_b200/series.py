"""Hermite (or Fourier) series: a coefficient vector in index-set order attached to a Position and a
weight factor (reference: hermipy/series.py:33-298).  All arithmetic on index maps (`project`, the
contraction behind `*`) goes through hermipy_b200.core, i.e. the CUDA library."""
import numpy as np
import numpy.linalg as la

from . import core
from . import function as func
from . import lib
from . import position as pos

very_small = 1e-10
_NUMBER = (int, float, np.integer, np.floating)


class Series:

    @staticmethod
    def tensorize(args):
        """Left fold of pairwise products.  Directions shared by two factors are contracted (weighted
        L2 inner product, core.inner); the others are tensorised."""
        def pair(s1, s2):
            if not isinstance(s1, Series) or not isinstance(s2, Series):
                raise ValueError("Arguments must be Series!")
            for scalar, other in ((s1, s2), (s2, s1)):
                if scalar.position.dim == 0:
                    return Series(scalar.coeffs[0] * other.coeffs, other.position, factor=other.factor,
                                  index_set=other.index_set)
            common = sorted(set(s1.position.dirs) & set(s2.position.dirs))
            w1, w2 = s1.factor.project(common), s2.factor.project(common)
            if not w1 == w2:                      # the contraction only makes sense in one weighted space
                raise ValueError("Factors should be match!")
            if s1.degree != s2.degree:
                raise ValueError("s1 and s2 should have the same degree")
            if s1.index_set != s2.index_set:
                raise ValueError("s1 and s2 should have the same index_set")
            coeffs = core.inner(s1.coeffs, s2.coeffs, s1.position.dirs, s2.position.dirs, index_set=s1.index_set)
            position = pos.Position.tensorize([s1.position, s2.position])
            factor = func.Function((s1.factor.sym / w1.sym) * (s2.factor.sym / w2.sym), dirs=position.dirs)
            return Series(coeffs, position, factor=factor, index_set=s1.index_set)

        result = args[0]
        for a in args[1:]:
            result = pair(result, a)
        return result

    def __init__(self, coeffs, position, factor=1, index_set="triangle", significant=0):
        self.coeffs = coeffs if isinstance(coeffs, np.ndarray) else np.asarray(coeffs)
        self.position = position
        self.index_set = index_set
        if position.dirs == []:
            self.degree = 0
        else:
            self.degree = core.iterator_get_degree(position.dim, len(coeffs), index_set=index_set)
            if len(self.multi_indices()) != len(coeffs):
                raise ValueError("Invalid arguments: length of self.coeffs does not match number of multi-indices!")
        if significant != 0:
            for i, c in enumerate(self.coeffs):
                self.coeffs[i] = round(c, significant)
        self.factor = func.Function(factor, dirs=self.position.dirs)
        if not self.coeffs.flags['C_CONTIGUOUS']:
            self.coeffs = self.coeffs.copy(order='C')

    def _like(self, coeffs):
        return Series(coeffs, self.position, factor=self.factor, index_set=self.index_set)

    def __eq__(self, other):
        if not isinstance(other, Series):
            raise ValueError("Invalid argument!")
        return self.position == other.position and self.factor == other.factor \
            and la.norm(self.coeffs - other.coeffs) < very_small

    def __add__(self, other):
        if isinstance(other, _NUMBER):
            return self._like(self.coeffs + other)
        if isinstance(other, Series):
            if not (self.position == other.position and self.index_set == other.index_set
                    and self.factor == other.factor):
                raise ValueError("Invalid argument!")
            return self._like(self.coeffs + other.coeffs)
        raise TypeError("Invalid type!)")

    def __mul__(self, other):
        if isinstance(other, _NUMBER):
            return self._like(self.coeffs * other)
        if isinstance(other, Series):
            if self.index_set != other.index_set:
                raise ValueError("Index sets must match!")
            return Series.tensorize([self, other])
        raise TypeError("Invalid type: " + str(type(other)))

    __rmul__ = __mul__

    def __sub__(self, other):
        return self + other * (-1)

    def __neg__(self):
        return self * (-1)

    def __truediv__(self, other):
        if not isinstance(other, _NUMBER):
            raise ValueError("Invalid argument!")
        return self._like(self.coeffs / other)

    def __repr__(self):
        indices = self.multi_indices()
        if len(indices) != len(self.coeffs):
            raise ValueError("Internal error: length of self.coeffs does not match number of multi-indices!")
        return "".join(str(m) + ": " + str(c) + "\n" for m, c in zip(indices, self.coeffs))

    def __float__(self):
        if self.position.dim != 0:
            raise ValueError("Dimension must be 0!")
        return float(self.coeffs[0])

    def __getitem__(self, index):
        return self.coeffs[index]

    def multi_indices(self):
        if self.position.dim == 0:
            return [None]
        return core.iterator_list_indices(self.position.dim, self.degree, index_set=self.index_set)

    def project(self, directions):
        if not isinstance(directions, list):
            directions = [directions]
        rel = [self.position.dirs.index(d) for d in directions]
        coeffs = core.project(self.coeffs, self.position.dim, rel, index_set=self.index_set)
        return Series(coeffs, self.position.project(directions), factor=self.factor.project(directions),
                      index_set=self.index_set)

    def subdegree(self, degree):
        """Truncate (or zero-extend) to another degree; relies on the enumeration being nested in the
        degree, which cross_nc is not."""
        if self.index_set == "cross_nc":
            raise ValueError("cross_nc not supported!")
        n_polys = core.iterator_size(self.position.dim, degree, index_set=self.index_set)
        if degree <= self.degree:
            coeffs = self.coeffs[0:n_polys]
        else:
            coeffs = np.hstack([self.coeffs, np.zeros(n_polys - len(self.coeffs))])
        return self._like(coeffs)

    def to_cross(self, degree):
        if self.index_set != "triangle" or degree + self.position.dim - 1 > self.degree:
            raise ValueError("Invalid argument!")
        coeffs = self.coeffs[lib.cross_in_triangle(self.position.dim, degree)]
        return Series(coeffs, self.position, index_set="cross")

    def to_function(self):
        """The series as a symbolic Function (un-weighted: the factor is not included)."""
        if not self.position.is_diag:
            raise ValueError("Position must be diagonal!")
        dirs = self.position.dirs
        basis = []
        for i, d in enumerate(dirs):
            # NOTE the reference scales by sqrt(cov) here (series.py:277-278); kept as is for parity
            mean, var = self.position.mean[i], self.position.cov[i][i]
            h = [func.Function('1', dirs=dirs), func.Function(np.sqrt(var) * (func.Function.xyz[d] - mean), dirs=dirs)]
            for k in range(1, self.degree):
                a, b = float(1 / np.sqrt(k + 1)), -float(np.sqrt(k) / np.sqrt(k + 1))
                h.append(h[-1] * h[1] * a + h[-2] * b)
            basis.append(h)
        total = func.Function('0', dirs=dirs)
        for c, m in zip(self.coeffs, self.multi_indices()):
            term = func.Function(c, dirs=dirs)
            for i in range(self.position.dim):
                term = term * basis[i][m[i]]
            total = total + term
        return total

    def plot(self, ax=None, title=None):
        from ._plotting import plot_series
        return plot_series(self, ax=ax, title=title)
