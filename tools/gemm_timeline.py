"""Developer script (needs a library built with `make -C hermipy_b200/csrc NVCC="nvcc -DHB_GEMM_TRACE"`): per-CTA timeline of every GEMM launch of one C2 forward + inverse transform (batch 17).
For each launch: when CTAs start relative to the first one, how long until their first operands land, main-loop
and epilogue durations, and what is left between the last CTA's end and ... the next launch's first CTA."""
import ctypes as C, os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan
from hermipy_b200 import _lib

lib = C.CDLL(_lib.LIB_PATH)
n, deg, B = 401, 200, int(sys.argv[1]) if len(sys.argv) > 1 else 17
x, w = sp.roots_hermitenorm(n); w = w / np.sqrt(2 * np.pi)
pl = Plan(deg, [x, x], [w, w], index_set="cube")
F = torch.randn(B, pl.grid_size, dtype=torch.float64, device="cuda")
Cf, G = pl.empty_coef(B), pl.empty_grid(B)
for _ in range(3):
    pl.forward(F, Cf); pl.backward(Cf, G)
torch.cuda.synchronize()
cap = 4096
us = lambda a: np.asarray(a) * 1e-3


def traced(fn, label):
    """run fn with one trace buffer per GEMM launch: the library writes into whatever buffer is set at launch time,
    so each launch gets its own by running the stages one at a time is not possible from here -- instead the CTAs of
    consecutive launches are separated by their start times (launches of one stream do not overlap)."""
    tr = torch.zeros(2 + 5 * cap, dtype=torch.int64, device="cuda"); tr[1] = cap
    lib.hb_debug_trace(C.c_void_p(tr.data_ptr()))
    pl.set_timing(True)
    fn(); torch.cuda.synchronize()
    timings = pl.timings()
    pl.set_timing(False)
    lib.hb_debug_trace(None)
    return tr.cpu().numpy(), timings


# the trace is indexed by blockIdx, so two launches in one call would overwrite each other: trace one launch at a
# time by toggling the buffer between two calls that differ in which stage is traced (HB_TRACE_TAG selects it)
for direction, fn in (("fwd", lambda: pl.forward(F, Cf)), ("bwd", lambda: pl.backward(Cf, G))):
    for tag in ((2, 1) if direction == "fwd" else (12, 11)):
        os.environ["HB_TRACE_TAG"] = str(tag)
        t, timings = traced(fn, direction)
        k = int(t[0])
        if k == 0:
            print(direction, "tag", tag, ": no records"); continue
        rec = t[2:2 + 5 * min(k, cap)].reshape(-1, 5)
        rec = rec[rec[:, 1] > 0]
        t0 = rec[:, 1].min()
        start, first, main, epi = us(rec[:, 1] - t0), us(rec[:, 2] - rec[:, 1]), us(rec[:, 3] - rec[:, 2]), us(rec[:, 4] - rec[:, 3])
        span = us(rec[:, 4].max() - t0)
        ev = [ms for tg, ms, _ in timings if tg == tag]
        per_sm = {}
        for r in rec:
            per_sm.setdefault(int(r[0]), []).append(r)
        print("%s tag %d: %d CTAs on %d SMs, event time %.1f us, first CTA start -> last CTA end %.1f us" % (
            direction, tag, len(rec), len(per_sm), ev[0] * 1e3 if ev else -1, span))
        for name, v in (("CTA start after first", start), ("start -> first operands", first), ("main loop", main), ("epilogue", epi),
                        ("CTA end after first start", us(rec[:, 4] - t0))):
            print("   %-26s us: mean %6.1f  p10 %6.1f  p50 %6.1f  p90 %6.1f  max %6.1f" % (name, v.mean(), *np.percentile(v, [10, 50, 90]), v.max()))
