"""Developer script: forward + inverse 2-D transform (401 x 401, degree 200 cube) of `batch` functions, graph replay,
with the parity split off (HB_TRANSFORM_PARITY=0) and forced (=2): where the split starts to pay off."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import gauss_hermite
    from hermipy_b200.plan import Plan
    n, deg = int(sys.argv[2]), int(sys.argv[3])
    x, w = gauss_hermite(n)
    plan = Plan(deg, [x, x], [w, w], index_set="cube")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    out = []
    for B in (1, 2, 4, 8, 12, 17, 32):
        F = torch.randn(B, plan.grid_size, dtype=torch.float64, device="cuda")
        C = plan.empty_coef(B); G = plan.empty_grid(B)
        for _ in range(3):
            plan.forward(F, C); plan.backward(C, G)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            plan.forward(F, C); plan.backward(C, G)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        gflop = 2 * B * (n * n * (deg + 1) + (deg + 1) ** 2 * n) * 1e-9
        out.append("B %2d (%.2f GFLOP/dir): %.1f us" % (B, gflop, float(np.median(ts))))
    print("PARITY=%s n %d degree %d: " % (os.environ["HB_TRANSFORM_PARITY"], n, deg) + "; ".join(out))
else:
    for n, deg in ((401, 200), (201, 100)):
        for p in ("0", "2"):
            subprocess.run([sys.executable, __file__, "run", str(n), str(deg)], env=dict(os.environ, HB_TRANSFORM_PARITY=p))
