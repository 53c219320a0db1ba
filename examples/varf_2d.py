"""2-D Galerkin matrix of a McKean-Vlasov-type generator on the triangle set, dense and sparse, with the separable
terms assembled from 1-D pieces (settings['tensorize']) -- the 2-D usage pattern of the reference's examples.

    python examples/varf_2d.py [degree]            (needs a B200)
"""
import os
import sys
import time

import numpy as np
import sympy as sym

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermipy as hm  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 40
x, y, f = hm.x, hm.y, hm.f
op = sym.diff(f(x, y), x, x) / 2 + sym.diff((x ** 3 - x + y / 2) * f(x, y), x) \
    + sym.diff(f(x, y), y, y) + sym.diff(y * f(x, y), y)
quad = hm.Quad.gauss_hermite(2 * degree + 1, dim=2)
for sparse in (False, True):
    t0 = time.time()
    mat = quad.discretize_op(op, degree, sparse=sparse).matrix
    dt = time.time() - t0
    nnz = mat.nnz if sparse else int(np.count_nonzero(mat))
    print("degree %d, %s: %d x %d, %d nonzeros, %.2f s" % (degree, "sparse" if sparse else "dense", *mat.shape, nnz, dt))
series = quad.transform(sym.exp(-(x * x + x * y + y * y) / 10), degree)
print("leading coefficients:", np.round(series.coeffs[:6], 6))
