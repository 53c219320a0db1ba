"""Developer script: Gram varf assembly times (C2 cube / triangle, C3 triangle) under the current environment.
    HB_VARF_FAMILIES=0 HB_VARF_GROUP=4 python tools/varf_variants.py [c3]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gauss_hermite, bistable
from hermipy_b200.plan import Plan

def run(n_rule, degree, set_name, reps=3):
    x, w = gauss_hermite(n_rule)
    keep = w > 0
    x, w = x[keep], w[keep]
    plan = Plan(degree, [x, x], [w, w], index_set=set_name)
    X, Y = np.meshgrid(x, x, indexing="ij")
    f = plan.pack_grid(np.exp(-3 * bistable(X, Y) / 10)[None])[0]
    out = torch.empty((plan.n_polys, plan.n_polys), dtype=torch.float64, device="cuda")
    plan.varf(f, "gram", out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.varf(f, "gram", out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%s degree %d %s: %.2f ms (min %.2f)" % (" ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("HB_VARF")),
                                                   degree, set_name, np.median(ts), min(ts)), flush=True)
    del out

run(401, 200, "cube")
run(401, 200, "triangle")
if len(sys.argv) > 1 and sys.argv[1] == "c3":
    run(801, 400, "triangle", reps=2)
