"""Developer script: the C2 transform step (batch 17, 401 x 401, degree 200) with and without the parity split."""
import os, sys
import numpy as np, scipy.special as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan

x, w = sp.roots_hermitenorm(401); w = w / np.sqrt(2 * np.pi)
x, w = (x - x[::-1]) / 2, (w + w[::-1]) / 2
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for s in ("cube", "triangle"):
    pl = Plan(200, [x, x], [w, w], index_set=s)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    f = torch.randn(B, pl.grid_size, dtype=torch.float64, device="cuda"); c = pl.empty_coef(B); g = pl.empty_grid(B)
    res = {}
    for mode in ("0", "2"):
        os.environ["HB_TRANSFORM_PARITY"] = mode
        def step():
            pl.forward(f, c); pl.backward(c, g)
        for _ in range(3): step()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        pl.set_timing(True)
        pl.forward(f, c); torch.cuda.synchronize(); lf = [(t, round(ms * 1e3, 1)) for t, ms, _ in pl.timings()]
        pl.backward(c, g); torch.cuda.synchronize(); lb = [(t, round(ms * 1e3, 1)) for t, ms, _ in pl.timings()]
        pl.set_timing(False)
        res[mode] = c.clone()
        print("%s parity=%s: step %.1f us (min %.1f); fwd GEMMs %s bwd GEMMs %s" % (s, mode, 1e3 * np.median(ts), 1e3 * min(ts), lf, lb))
    print("   coefficient rel diff parity vs plain: %.2e" % float((res["0"] - res["2"]).norm() / res["0"].norm()))
