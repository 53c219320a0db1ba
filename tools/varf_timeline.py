"""Developer script: per-tile timeline of the persistent varf stage-2 kernel (hb_varf_gemm.cu): for every tile the
times at which its group started waiting, entered / left the main loop, got the staging buffer and finished."""
import ctypes as C, os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan
from hermipy_b200 import _lib

lib = C.CDLL(_lib.LIB_PATH)
x, w = sp.roots_hermitenorm(401); w = w / np.sqrt(2 * np.pi)
X, Y = np.meshgrid(x, x, indexing="ij")
V = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
pl = Plan(200, [x, x], [w, w], index_set=sys.argv[1] if len(sys.argv) > 1 else "cube")
out = torch.empty((pl.n_polys, pl.n_polys), dtype=torch.float64, device="cuda")
fE = pl.pack_grid(np.exp(-3 * V / 10)[None])[0]
pl.varf(fE, "gram", out); torch.cuda.synchronize()
cap = 200000
tr = torch.zeros(2 + 10 * cap, dtype=torch.int64, device="cuda"); tr[1] = cap
lib.hb_debug_trace(C.c_void_p(tr.data_ptr()))
pl.varf(fE, "gram", out); torch.cuda.synchronize()
lib.hb_debug_trace(None)
t = tr.cpu().numpy(); n = int(min(t[0], cap)); rec = t[2:2 + 10 * n].reshape(n, 10)
print("tiles", n)
t0 = rec[:, 2].min()
us = lambda a: a * 1e-3
wait, main, hand, epi = us(rec[:, 3] - rec[:, 2]), us(rec[:, 4] - rec[:, 3]), us(rec[:, 5] - rec[:, 4]), us(rec[:, 6] - rec[:, 5])
stg, dire, mirr = us(rec[:, 7] - rec[:, 5]), us(rec[:, 8] - rec[:, 7]), us(rec[:, 6] - rec[:, 8])
for name, v in (("wait for turn", wait), ("main loop", main), ("wait for staging", hand), ("stage+scatter", epi), ("  staging", stg), ("  direct pass", dire), ("  mirror pass", mirr)):
    print("%-18s us: mean %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f" % (name, v.mean(), *np.percentile(v, [10, 50, 90]), v.max()))
print("kernel span ms %.2f" % ((rec[:, 6].max() - t0) * 1e-6))
ends = {}
for r in rec:
    ends[r[1]] = max(ends.get(r[1], 0), r[6])
e = np.array(list(ends.values())); print("group finish times ms: min %.2f max %.2f" % ((e.min() - t0) * 1e-6, (e.max() - t0) * 1e-6))
for grp in (0, 1, 100, 101):
    r = rec[rec[:, 1] == grp]; r = r[np.argsort(r[:, 2])][:5]
    print("group %d first tiles (wait, main0, main1, staged, end) us:" % grp, [tuple(round((v - t0) * 1e-3, 1) for v in row[2:7]) for row in r])
