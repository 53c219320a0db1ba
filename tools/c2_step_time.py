"""Developer script: the C2 transform step (17 functions, 401 x 401, degree 200 cube) forward / inverse times with
per-launch breakdown.   HB_TRANSFORM_FLATTEN=0 python tools/c2_step_time.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gauss_hermite
from hermipy_b200.plan import Plan
x, w = gauss_hermite(401)
plan = Plan(200, [x, x], [w, w], index_set="cube")
B = 17
F = torch.randn(B, plan.grid_size, dtype=torch.float64, device="cuda")
C = plan.empty_coef(B); G = plan.empty_grid(B)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for _ in range(3):
    plan.forward(F, C); plan.backward(C, G)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    plan.forward(F, C); plan.backward(C, G)
ts = []
for _ in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
plan.set_timing(True)
acc = {}
for _ in range(10):
    flush.zero_()
    plan.forward(F, C)
    for tag, ms, _ in plan.timings(): acc.setdefault(("f", tag), []).append(ms * 1e3)
    plan.backward(C, G)
    for tag, ms, _ in plan.timings(): acc.setdefault(("b", tag), []).append(ms * 1e3)
print("FLATTEN=%s: step %.1f us median (min %.1f); launches %s" % (os.environ.get("HB_TRANSFORM_FLATTEN", "1"), np.median(ts), min(ts),
      {k: round(float(np.mean(v)), 1) for k, v in acc.items()}))
