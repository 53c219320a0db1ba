"""Spectral gap of a 2-D Fokker-Planck generator without the matrix ever leaving the GPU.

    python examples/resident_generator.py [degree]

`discretize_op(..., resident=True)` assembles the Galerkin matrix in HBM; `eigs` runs ARPACK on the host with
the matrix-vector product (`hb_dev_matvec`, HBM-bound) on the device, so only vectors cross PCIe.  At
degree 200 on the cube set the matrix is 13 GB: one product takes 2 ms.
"""
import os
import sys

import numpy as np
import sympy as sym

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermipy_b200 as hm  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 40
x, y, f = hm.x, hm.y, hm.f
V = x ** 4 / 4 - x ** 2 / 2 + y ** 2 / 2 + x * y / 2
u = f(x, y)
generator = sym.diff(sym.diff(V, x) * u + sym.diff(u, x), x) + sym.diff(sym.diff(V, y) * u + sym.diff(u, y), y)

quad = hm.Quad.gauss_hermite(2 * degree + 1, dim=2, cov=[[0.4, 0], [0, 0.8]], factor=sym.exp(-V / 2))
A = quad.discretize_op(generator, degree, index_set="triangle", resident=True)
print("matrix: %d x %d, %.2f GB in HBM" % (A.matrix.n, A.matrix.n, 8e-9 * A.matrix.n ** 2))
values, vectors = A.eigs(k=3, sigma=0.5)
order = np.argsort(-values.real)
print("leading eigenvalues:", values.real[order])
print("spectral gap: %.6f" % -values.real[order][1])
