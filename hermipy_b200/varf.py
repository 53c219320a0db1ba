"""Galerkin matrix of an operator in the Hermite basis (reference: hermipy/varf.py:36-284): a dense
ndarray or scipy CSR matrix in index-set order attached to a Position and a weight factor.  The matrix
comes from the CUDA assembly (core.varf / core.varfd / core.tensorize); the linear algebra on it
(`solve`, `eigs`) is SciPy's, as in the reference."""
import numpy as np
import numpy.linalg as la
import scipy.sparse as ss
import scipy.sparse.linalg as spla

from . import core
from .cache import cache as _memoise
from . import function as func
from . import lib
from . import position as pos
from . import series as hs
from .settings import settings
from . import stats

very_small = 1e-10
_NUMBER = (int, float, np.integer, np.floating)


class Varf:

    @staticmethod
    def tensorize(args, sparse=None):
        """Kronecker product, restricted to the index set, of matrices living on disjoint directions."""
        if any(not isinstance(a, Varf) for a in args):
            raise ValueError("Invalid argument(s)")
        if len(args) == 1:
            return args[0]
        sparse = settings['sparse'] if sparse is None else sparse
        first = args[0]
        if any(a.index_set != first.index_set or a.degree != first.degree for a in args):
            raise ValueError("Invalid argument(s)")
        matrices = {frozenset(a.position.dirs): a.matrix for a in args}
        matrix = core.tensorize(matrices, sparse=sparse, index_set=first.index_set)
        position = pos.Position.tensorize([a.position for a in args])
        weight = 1
        for a in args:
            weight = weight * a.factor.sym
        return Varf(matrix, position, factor=func.Function(weight, dirs=position.dirs), index_set=first.index_set)

    def __init__(self, matrix, position, factor=1, index_set="triangle"):
        self.is_sparse = isinstance(matrix, ss.csr_matrix)
        self.matrix = matrix
        self.position = position
        self.index_set = index_set
        self.degree = core.iterator_get_degree(position.dim, matrix.shape[0], index_set=index_set)
        assert len(self.multi_indices()) == matrix.shape[0]
        self.factor = func.Function(factor, dirs=position.dirs)

    def _like(self, matrix):
        return Varf(matrix, self.position, factor=self.factor, index_set=self.index_set)

    def __eq__(self, other):
        if not isinstance(other, Varf):
            raise ValueError("Invalid argument!")
        norm = spla.norm if self.is_sparse and other.is_sparse else la.norm
        return self.position == other.position and norm(self.matrix - other.matrix) < very_small

    def __add__(self, other):
        if isinstance(other, _NUMBER):
            return self._like(self.matrix + other)
        if isinstance(other, Varf):
            if not (self.position == other.position and self.index_set == other.index_set
                    and self.factor == other.factor):
                raise ValueError("Invalid argument!")
            return self._like(self.matrix + other.matrix)
        raise TypeError("Invalid type!)")

    def __mul__(self, other):
        if isinstance(other, _NUMBER):
            return self._like(self.matrix * other)
        if isinstance(other, Varf):
            if self.index_set != other.index_set:
                raise ValueError("Invalid argument!")
            return Varf.tensorize([self, other])
        raise TypeError("Invalid type: " + str(type(other)))

    __rmul__ = __mul__
    __radd__ = __add__

    def __sub__(self, other):
        return self + (other * (-1))

    def __neg__(self):
        return self * (-1)

    def __call__(self, series):
        """Apply the operator to a Series (matrix-vector product)."""
        if not isinstance(series, hs.Series) or not self.position == series.position \
                or self.index_set != series.index_set:
            raise ValueError("Invalid argument!")
        return hs.Series(self.matrix.dot(series.coeffs), self.position, factor=self.factor, index_set=self.index_set)

    def __getitem__(self, index):
        return self.matrix[index]

    def multi_indices(self):
        return core.iterator_list_indices(self.position.dim, self.degree, index_set=self.index_set)

    def project(self, directions):
        if isinstance(directions, int):
            directions = [directions]
        rel = [self.position.dirs.index(d) for d in directions]
        matrix = core.project(self.matrix, self.position.dim, rel, index_set=self.index_set)
        return Varf(matrix, self.position.project(directions), factor=self.factor.project(directions),
                    index_set=self.index_set)

    def subdegree(self, degree):
        if degree > self.degree:
            raise ValueError("Invalid argument: degree too high!")
        n = core.iterator_size(self.position.dim, degree, index_set=self.index_set)
        block = self.matrix[0:n, 0:n]
        return self._like(block.copy() if self.is_sparse else block.copy(order='C'))

    def to_cross(self, degree):
        if self.index_set != "triangle" or degree + self.position.dim - 1 > self.degree:
            raise ValueError("Invalid argument!")
        keep = lib.cross_in_triangle(self.position.dim, degree)
        return Varf(self.matrix[np.ix_(keep, keep)], self.position, factor=self.factor, index_set="cross")

    def solve(self, series, use_gmres=False, remove0=False, remove_vec=None, preconditioner=None, quiet=False,
              **kwargs):
        """Solve A u = series.  `remove0` / `remove_vec` border the system with a vector (the null space
        found by eigs, or the one given) and a Lagrange multiplier so that a singular operator (e.g. a
        Fokker-Planck generator) becomes solvable."""
        if not self.position == series.position or self.index_set != series.index_set:
            raise ValueError("Invalid argument: position and index set must coincide!")
        bordered = remove0 or remove_vec is not None
        if bordered:
            if remove0:
                [e], [remove_vec] = self.eigs(k=1, which='SM')
                if not quiet:
                    print("Removing eigenspace with eigenvalue {}.".format(abs(e)))
            else:
                remove_vec = remove_vec / np.sqrt(float(remove_vec * remove_vec))
                e = float(remove_vec * self(remove_vec))
                if not quiet:
                    print("Removing vector with value {}.".format(abs(e)))
            if abs(e) > 0.01:
                print("[Hermipy:varf:solve warning] Value of e: {}".format(e))
            vstack, hstack = (ss.vstack, ss.hstack) if self.is_sparse else (np.vstack, np.hstack)
            border = np.asarray(remove_vec.coeffs)
            matrix = vstack((self.matrix, border))
            matrix = hstack((matrix, np.append(border, 0).reshape(-1, 1)))
            vector = np.append(series.coeffs, 0)
            if self.is_sparse:
                matrix = matrix.tocsr()
        else:
            matrix, vector = self.matrix, series.coeffs

        if use_gmres:
            M = None
            if preconditioner == 'ilu':
                ilu = spla.spilu(matrix)
                M = spla.LinearOperator(matrix.shape, ilu.solve)
            # callbacks and operators are not hashable content: they do not enter the cache key
            gmres = _memoise(quiet=True, hash_extend=lambda arg: '0')(spla.gmres)
            solution = gmres(matrix, vector, M=M, **kwargs)[0]
        else:
            direct = _memoise(quiet=True)(spla.spsolve if self.is_sparse else la.solve)
            solution = direct(matrix, vector, **kwargs)
        if bordered:
            solution = np.array(solution[0:-1])
        return hs.Series(solution, position=self.position, factor=self.factor, index_set=self.index_set)

    @stats.log_stats()
    def eigs(self, **kwargs):
        """scipy.sparse.linalg.eigs on the matrix; eigenvectors come back as Series (real when they are)."""
        values, vectors = _memoise(quiet=True)(spla.eigs)(self.matrix, **kwargs)
        series = []
        for v in vectors.T:
            v = v if la.norm(np.imag(v)) > 1e-15 else np.real(v)
            series.append(hs.Series(v, self.position, factor=self.factor, index_set=self.index_set))
        return values, series

    def plot(self, ax=None, lines=True):
        from ._plotting import plot_varf
        return plot_varf(self, ax=ax, lines=lines)
