"""Developer script: many small Gram varf assemblies with random degrees, index sets and owned-row ranges, to shake
out scheduling problems of the persistent stage-2 kernel (tile counts of 1, 2, odd, fewer than the groups ...).
Every result is compared with the full matrix of the same plan."""
import os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for case in range(n_cases):
    degree = int(rng.integers(1, 40))
    n = int(rng.integers(max(4, degree + 1), 2 * degree + 12))
    index_set = ["triangle", "cube", "rectangle", "cross", "cross_nc"][int(rng.integers(0, 5))]
    x, w = sp.roots_hermitenorm(n); w = w / np.sqrt(2 * np.pi)
    if rng.random() < 0.3:
        x = x + 0.1          # not symmetric: no parity split
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = np.exp(-(X ** 2 + Y ** 2) / 9) * (1 + np.sin(X + 0.3 * Y))
    try:
        pl = Plan(degree, [x, x], [w, w], index_set=index_set)
    except Exception as e:      # cross with degree < 2 etc. is rejected like the reference does
        print("case %d skipped (%s)" % (case, str(e)[:60])); continue
    fd = pl.pack_grid(f[None])[0]
    full = pl.varf(fd, "gram")
    N = pl.n_polys
    for _ in range(4):
        lo = int(rng.integers(0, N)); hi = int(rng.integers(lo + 1, N + 1))
        part = pl.varf(fd, "gram", row_begin=lo, row_end=hi)
        assert torch.equal(part, full[lo:hi]), (case, degree, n, index_set, lo, hi)
    assert torch.equal(full, full.T)
    torch.cuda.synchronize()
    if case % 10 == 0:
        print("case %d ok: degree %d n %d %s n_polys %d" % (case, degree, n, index_set, N), flush=True)
print("all %d cases ok" % n_cases)
