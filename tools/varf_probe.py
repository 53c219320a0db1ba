"""Developer script: time the C2 dense (Gram) varf stage-2 GEMM, cube and triangle, optionally under HB_VARF_DEBUG."""
import os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan

x, w = sp.roots_hermitenorm(401); w = w / np.sqrt(2 * np.pi)
X, Y = np.meshgrid(x, x, indexing="ij")
V = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
for s in sys.argv[1:] or ["cube", "triangle"]:
    pl = Plan(200, [x, x], [w, w], index_set=s)
    pad = int(os.environ.get("PITCH_ALIGN", "1"))
    pitch = (pl.n_polys + pad - 1) // pad * pad
    out = torch.empty((pl.n_polys, pitch), dtype=torch.float64, device="cuda")[:, :pl.n_polys]
    fE = pl.pack_grid(np.exp(-3 * V / 10)[None])[0]
    pl.varf(fE, "gram", out); torch.cuda.synchronize()
    pl.set_timing(True)
    ts = []
    for _ in range(3):
        pl.varf(fE, "gram", out); torch.cuda.synchronize()
        ts.append({t: ms for t, ms, _ in pl.timings()}.get(22))
    print(s, "pitch", out.stride(0), "HB_VARF_DEBUG=%s" % os.environ.get("HB_VARF_DEBUG", ""), "stage2 ms", ["%.2f" % t for t in ts], "sym err", float((out - out.T).abs().max()))
    del out, pl
