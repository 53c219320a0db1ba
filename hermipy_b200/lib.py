"""Counterpart of the reference's hermipy/lib.py:24-33 (`hermegauss_nd`): Gauss-Hermite nodes and
weights for the weight exp(-x^2/2)/sqrt(2 pi).  The reference calls numpy.polynomial.hermite_e.hermegauss,
which returns zero/NaN weights for n >= 371 with NumPy 2.3 (SURVEY.md finding 3); this generator works
for any n: Golub-Welsch eigenvalues, Newton polishing on the scaled three-term recurrence, and weights
w_i = 1 / (n phi_{n-1}(x_i)^2) formed in the log domain (weights below 1e-308 underflow to exactly 0,
which is what they are in double precision).  Host-side setup code: O(n^2) flops, not on the hot path.
"""
import numpy as np
import scipy.linalg as sl


def _scaled_recurrence(x, n):
    """phi_{n-1}(x), phi_n(x) for the normalised polynomials phi_k = He_k/sqrt(k!), as (mantissas, log-scale):
    phi_k = mant_k * exp(logs), with the pair rescaled whenever it grows."""
    pm, p = np.ones_like(x), x.copy()        # phi_0, phi_1
    logs = np.zeros_like(x)
    if n == 0:
        return np.zeros_like(x), pm, logs
    for k in range(1, n):
        a, b = 1 / np.sqrt(k + 1.0), np.sqrt(k / (k + 1.0))
        pm, p = p, a * x * p - b * pm
        big = np.maximum(np.abs(p), np.abs(pm))
        scale = np.where(big > 1e100, 1e-100, 1.0)
        pm, p = pm * scale, p * scale
        logs -= np.log(scale)
    return pm, p, logs


def hermegauss(n, sqrt_weights=False):
    """n-point rule: nodes (ascending) and weights normalised to sum 1.  sqrt_weights=True also returns sqrt(w_i)
    formed in the log domain: representable while w_i > 1e-616, where w_i itself underflows below 1e-308 (the Gram-form
    varf at high degree needs it: Plan(..., sqrt_weights=...))."""
    n = int(n)
    if n < 1:
        raise ValueError("n must be positive")
    if n == 1:
        return (np.zeros(1), np.ones(1), np.ones(1)) if sqrt_weights else (np.zeros(1), np.ones(1))
    x = sl.eigvalsh_tridiagonal(np.zeros(n), np.sqrt(np.arange(1.0, n)))
    x = (x - x[::-1]) / 2                      # enforce the exact symmetry of the rule
    for _ in range(3):                         # Newton: phi_n' = sqrt(n) phi_{n-1}
        pm, p, _ = _scaled_recurrence(x, n)
        x = x - p / (np.sqrt(n) * pm)
        x = (x - x[::-1]) / 2
    pm, _, logs = _scaled_recurrence(x, n)
    with np.errstate(divide="ignore", under="ignore"):
        logw = -np.log(float(n)) - 2 * (np.log(np.abs(pm)) + logs)
        w = np.exp(logw)
    w = (w + w[::-1]) / 2
    total = w.sum()
    if not sqrt_weights:
        return x, w / total
    with np.errstate(under="ignore"):
        sw = np.exp(0.5 * logw)
    sw = (sw + sw[::-1]) / 2
    return x, w / total, sw / np.sqrt(total)


def gram_rule(n, degree, tol=1e-17):
    """The part of the n-point rule a degree-`degree` Gram-form varf needs: (nodes, weights, sqrt_weights) restricted
    to the nodes where some psi_k = sqrt(w) phi_k, k <= degree, exceeds `tol` (the functions decay like Gaussians beyond
    their turning point 2 sqrt(k)), symmetric bit for bit.  Setup code on the host."""
    x, w, sw = hermegauss(n, sqrt_weights=True)
    pm, p = sw.copy(), x * sw
    amp = np.maximum(np.abs(pm), np.abs(p))
    for k in range(1, int(degree)):
        pm, p = p, (x * p - np.sqrt(float(k)) * pm) / np.sqrt(k + 1.0)
        amp = np.maximum(amp, np.abs(p))
    keep = amp > tol
    keep &= keep[::-1]
    return x[keep], w[keep], sw[keep]


def hermegauss_nd(n_points):
    """hermipy/lib.py:24-33"""
    nodes, weights = [], []
    for n in n_points:
        x, w = hermegauss(n)
        nodes.append(x)
        weights.append(w)
    return nodes, weights


def cross_in_triangle(dim, degree):
    """Positions, in the triangle enumeration, of the multi-indices of the cross set (hermipy/lib.py:36-38)."""
    from . import core
    return [core.iterator_index(m) for m in core.iterator_list_indices(dim, degree, index_set="cross")]


def finest_common(set1, set2):
    """Finest partition that both partitions (sets of frozensets of directions) refine (hermipy/lib.py:41-68):
    the connected components of the 'shares a direction' relation over all blocks of both partitions."""
    blocks = [frozenset(b) for b in list(set1) + list(set2) if b]   # the empty block (constant factor) is not a group
    merged = []
    for block in blocks:
        touching = [m for m in merged if m & block]
        for m in touching:
            merged.remove(m)
            block = block | m
        merged.append(block)
    return set(merged)
