"""Where a quadrature / series lives: directions, mean, covariance and basis type per direction
(reference: hermipy/position.py:25-159).  The Hermite basis attached to a Position is the one orthonormal
for the Gaussian N(mean, cov); `factor` (= eigvec * sqrt(eigval) of cov) maps standard nodes to physical
nodes, x = mean + factor @ node.  For a Fourier direction (mean, cov) encode the interval instead:
mean = midpoint, cov = ((right-left)/2pi)^2 (quad.py:126-134)."""
import numpy as np
import numpy.linalg as la
import sympy as sym

from . import function as func


class Position:

    very_small = 1e-10

    @staticmethod
    def tensorize(args):
        """Combine positions on (mostly) disjoint directions.  A direction may appear in at most two
        arguments; a direction shared by two arguments must carry the same mean and variance in both and
        is contracted away (it does not appear in the result)."""
        count = {}
        for a in args:
            if not isinstance(a, Position) or not a.is_diag:
                raise ValueError("Invalid arguments!")
            for d in a.dirs:
                count[d] = count.get(d, 0) + 1
                if count[d] > 2:
                    raise ValueError("Direction appears more than once")

        def pair(p, q):
            for d in set(p.dirs) & set(q.dirs):
                i, j = p.dirs.index(d), q.dirs.index(d)
                if p.mean[i] != q.mean[j] or p.cov[i][i] != q.cov[j][j]:
                    raise ValueError("Mean or Cov do not match!")
            dirs = sorted(set(p.dirs) ^ set(q.dirs))
            mean, var, types = [], [], []
            for d in dirs:
                src = p if d in p.dirs else q
                k = src.dirs.index(d)
                mean.append(src.mean[k])
                var.append(src.cov[k][k])
                types.append(src.types[k])
            return Position(dirs=dirs, mean=mean, cov=np.diag(var), types=types)

        result = args[0]
        for a in args[1:]:
            result = pair(result, a)
        return result

    def __init__(self, dim=None, mean=None, cov=None, dirs=None, types=None):
        given = [len(v) for v in (mean, cov, dirs) if v is not None] + ([dim] if dim is not None else [])
        if not given:
            raise ValueError("All args are None!")
        self.dim = given[0]
        if any(g != self.dim for g in given) or (types is not None and len(types) != self.dim):
            raise ValueError("Invalid arguments dim/dirs")
        self.dirs = list(range(self.dim)) if dirs is None else dirs
        self.mean = np.zeros(self.dim) if mean is None else np.asarray(mean, float)
        self.cov = np.eye(self.dim) if cov is None else np.asarray(cov, float)
        self.types = ["hermite"] * self.dim if types is None else types
        if self.dim:
            eigval, eigvec = la.eig(self.cov)
            self.factor = np.matmul(eigvec, np.sqrt(np.diag(eigval)))
        else:
            self.factor = np.zeros((0, 0))
        off_diagonal = self.cov - np.diag(np.diag(self.cov))
        self.is_diag = not (self.dim > 0 and la.norm(off_diagonal, 2) > 1e-10)

    def __eq__(self, other):
        return self.dirs == other.dirs and self.types == other.types \
            and la.norm(self.mean - other.mean, 2) < self.very_small \
            and la.norm(self.cov - other.cov, 2) < self.very_small

    def __mul__(self, other):
        if not isinstance(other, Position):
            raise ValueError("Invalid argument!")
        return Position.tensorize([self, other])

    def __repr__(self):
        return "Dimension: {}\nDirections: {}\nMean: {}\nCovariance:\n{}\nTypes:\n{}".format(
            self.dim, self.dirs, self.mean, self.cov, self.types)

    def weight(self):
        """Gaussian density of the Hermite directions, as a SymPy expression in x, y, z, ..."""
        ind = [i for i, t in enumerate(self.types) if t == "hermite"]
        if not ind:
            return sym.Rational(1)
        mean, cov = self.mean[ind], self.cov[np.ix_(ind, ind)]
        delta = np.array([func.Function.xyz[self.dirs[i]] for i in ind]) - mean
        potential = 0.5 * la.inv(cov).dot(delta).dot(delta)
        normalization = 1 / sym.sqrt((2 * sym.pi) ** self.dim * la.det(cov))
        return normalization * sym.exp(-potential)

    def weights(self):
        if not self.is_diag:
            raise ValueError("Invalid: is_diag must be True")
        return [self.project(d).weight() for d in self.dirs]

    def project(self, directions):
        if isinstance(directions, int):
            directions = [directions]
        if not self.is_diag or directions != sorted(directions):
            raise ValueError("Invalid argument!")
        rel = [self.dirs.index(d) for d in directions]
        return Position(dim=len(rel), dirs=directions, mean=[self.mean[r] for r in rel],
                        cov=np.diag([self.cov[r][r] for r in rel]) if rel else np.zeros((0, 0)),
                        types=[self.types[r] for r in rel])
