"""Developer script (one GPU): the per-rank launch sequence of the sharded parity-split 3-D transform at world size
`g`, exchange kept local -- per-phase device times (events, eager launches) and the graph-replayed round trip.
    python tools/shard_emulate.py [g] [n]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gauss_hermite
from hermipy_b200.plan import Plan
from hermipy_b200.dist import ShardedParityTransform, plan_parity_tables, DeviceOps

g = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
x, w = gauss_hermite(n)
plan = Plan(n - 1, [x] * 3, [w] * 3, index_set="cube")
tables = plan_parity_tables(plan)


class TimedOps(DeviceOps):
    def __init__(self):
        self.marks, self.on = [], False

    def _mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def gemm(self, gm):
        self._mark("pre")
        super().gemm(gm)
        self._mark("gemm M%d N%d K%d b%dx%d%s" % (gm.M, gm.N, gm.K, gm.batch, gm.batch_lo, " scatter" if gm.blocks else ""))

    def parity_pass(self, merge, **kw):
        self._mark("pre")
        super().parity_pass(merge, **kw)
        self._mark("parity %s" % ("merge" if merge else "split"))


ops = TimedOps()
T = ShardedParityTransform([n] * 3, [n] * 3, tables, ops=ops, emulate=g)
f = torch.randn(T.local_grid_shape(), dtype=torch.float64, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = T.forward(f); T.backward(c)
torch.cuda.synchronize()
acc = {}
ops.on = True
for it in range(20):
    flush.zero_()
    ops.marks = []
    c = T.forward(f); T.backward(c)
    torch.cuda.synchronize()
    m = ops.marks
    for i in range(1, len(m), 2):
        acc.setdefault((i // 2, m[i][0]), []).append(m[i - 1][1].elapsed_time(m[i][1]) * 1e3)
ops.on = False
tot = 0
for (i, name), v in sorted(acc.items()):
    print("%2d %-40s %7.1f us (min %.1f)" % (i, name, np.median(v), min(v)))
    tot += np.median(v)
print("sum of phases %.1f us" % tot)
graph, c, gout = T.capture(f)
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("world %d emulated on one GPU, n %d: graph round trip %.1f us median, %.1f us min" % (g, n, np.median(ts), min(ts)))
