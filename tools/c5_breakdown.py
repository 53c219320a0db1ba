"""Developer script: per-launch times of the C5 batched 1-D transform (65536 x degree 1000) with the parity split."""
import os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan

x, w = sp.roots_hermitenorm(2001); w = w / np.sqrt(2 * np.pi)
keep = (w > 0) & (np.abs(x) <= 50); x, w = x[keep], w[keep]
pl = Plan(1000, [x], [w]); B = 65536
f = torch.randn(B, pl.grid_size, dtype=torch.float64, device="cuda"); c = pl.empty_coef(B); g = pl.empty_grid(B)
def timeit(fn, reps=5):
    for _ in range(2): fn()
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
    pl.forward(f, c); torch.cuda.synchronize(); lf = [round(ms, 3) for t, ms, _ in pl.timings()]
    pl.backward(c, g); torch.cuda.synchronize(); lb = [round(ms, 3) for t, ms, _ in pl.timings()]
    pl.set_timing(False)
    print("parity=%s fwd %.3f ms (GEMMs %s) bwd %.3f ms (GEMMs %s)" % (mode, tf, lf, tb, lb))
