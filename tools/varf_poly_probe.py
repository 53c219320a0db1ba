"""Developer script: time the reference-algebra (polynomial f) varf of the C2 configuration and its parts."""
import os, sys, time
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan

set_name = sys.argv[1] if len(sys.argv) > 1 else "cube"
x, w = sp.roots_hermitenorm(401); w = w / np.sqrt(2 * np.pi)
X, Y = np.meshgrid(x, x, indexing="ij")
V = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
pl = Plan(200, [x, x], [w, w], index_set=set_name)
out = torch.empty((pl.n_polys, pl.n_polys), dtype=torch.float64, device="cuda")
fV = pl.pack_grid(V[None])[0]


def timeit(fn, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    ts.sort()
    return ts[len(ts) // 2]


print("%s: varf(reference) event %.3f ms, wall %.3f ms" % ((set_name,) + timeit(lambda: pl.varf(fV, "reference", out))))
print("zero_ of the matrix: event %.3f ms, wall %.3f ms" % timeit(lambda: out.zero_()))
p2 = Plan(400, [x, x], [w, w], index_set=set_name)
c = p2.empty_coef(1)
print("transform at degree 2N: event %.3f ms, wall %.3f ms" % timeit(lambda: p2.forward(fV[None], c)))
