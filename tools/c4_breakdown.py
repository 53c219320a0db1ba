"""Developer script: per-launch times of the C4 (256^3) forward and inverse transform with and without the parity split."""
import os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan

x, w = sp.roots_hermitenorm(256); w = w / np.sqrt(2 * np.pi)
pl = Plan(255, [x] * 3, [w] * 3, index_set="cube")
f = torch.randn(1, pl.grid_size, dtype=torch.float64, device="cuda"); c = pl.empty_coef(1); g = pl.empty_grid(1)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mode in ("1", "0"):
    os.environ["HB_TRANSFORM_PARITY"] = mode
    tf, tb = timeit(lambda: pl.forward(f, c)), timeit(lambda: pl.backward(c, g))
    pl.set_timing(True)
    pl.forward(f, c); torch.cuda.synchronize(); lf = [(t, round(ms * 1e3, 1)) for t, ms, _ in pl.timings()]
    pl.backward(c, g); torch.cuda.synchronize(); lb = [(t, round(ms * 1e3, 1)) for t, ms, _ in pl.timings()]
    pl.set_timing(False)
    print("parity=%s fwd %.3f ms (GEMMs us %s = %.0f) bwd %.3f ms (GEMMs us %s = %.0f)" % (mode, tf, lf, sum(m for _, m in lf), tb, lb, sum(m for _, m in lb)))
