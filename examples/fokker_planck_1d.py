"""BASELINE.json configs[0] through the object layer: 1-D Fokker-Planck operator in a weighted Hermite basis,
degree 100 on 201 Gauss-Hermite nodes -- discretize_op (NVRTC discretize + varf / varfd kernels) -> eigs -> eval.

    python examples/fokker_planck_1d.py            (needs a B200; `import hermipy` resolves to hermipy_b200)
"""
import os
import sys
import time

import numpy as np
import scipy.sparse.linalg as las
import sympy as sym

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hermipy as hm  # noqa: E402

x, f = sym.symbols('x', real=True), sym.Function('f')
beta, m, s2, degree = sym.Rational(3), sym.Rational(1, 10), sym.Rational(1, 10), 100
Vp = x ** 4 / 4 - x ** 2 / 2                                   # bistable potential
Vq = sym.Rational(1, 2) * (x - m) ** 2 / (beta * s2)           # Gaussian the basis is attached to
forward = sym.diff(sym.diff(Vp, x) * f(x) + 1 / beta * sym.diff(f(x), x), x)
factor = sym.exp(-beta / 2 * (Vq + Vp))
backward = (forward.subs(f(x), f(x) * factor).doit() / factor).expand()

quad = hm.Quad.gauss_hermite(2 * degree + 1, dim=1, mean=[float(m)], cov=[[float(s2)]])
t0 = time.time()
operator = quad.discretize_op(backward, degree)                # Varf: 101 x 101 Galerkin matrix
t1 = time.time()
_, vecs = las.eigs(operator.matrix, k=1, which='LR')
ground = np.real(vecs.T[0])
ground *= np.sign(ground[0])
density = quad.eval(quad.series(ground)) * quad.discretize(factor)
density /= quad.norm(density, n=1, flat=True)
exact = quad.discretize(sym.exp(-beta * Vp))
exact /= quad.norm(exact, n=1, flat=True)
print("discretize_op: %.3f s (first call compiles the coefficient expressions with NVRTC)" % (t1 - t0))
print("invariant density, L2 error vs exp(-beta V): %.2e" % quad.norm(density - exact, flat=True))
series = quad.transform(sym.exp(-x * x / 8) * sym.cos(3 * x), degree)
print("transform -> eval round trip (weighted): %.2e"
      % np.linalg.norm((quad.eval(series) - quad.discretize(sym.exp(-x * x / 8) * sym.cos(3 * x))) * np.sqrt(quad.weights[0])))
hm.stats.print_stats() if hasattr(hm, "stats") else None
