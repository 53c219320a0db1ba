"""Developer script (torchrun): per-stage device times of the sharded 3-D transform (p2p exchange)."""
import os, sys
import numpy as np, torch, torch.distributed as dist, scipy.special as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan
from hermipy_b200 import dist as hd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256
x, w = sp.roots_hermitenorm(n); w = w / np.sqrt(2 * np.pi)
plan = Plan(n - 1, [x] * 3, [w] * 3, index_set="cube")
g = np.exp(-(x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2) / 8)
s = n // world
f = torch.from_numpy(np.ascontiguousarray(g[rank * s:(rank + 1) * s])).cuda()
T = hd.ShardedTransform([n] * 3, [n] * 3, hd.plan_tables_as_tensors(plan), exchange=sys.argv[1] if len(sys.argv) > 1 else "p2p")
marks = []
orig_gemm = T.gemm
def timed_gemm(A, B, out):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); orig_gemm(A, B, out); e1.record(); marks.append(("gemm %s" % (tuple(out.shape),), e0, e1))
T.gemm = timed_gemm
import hermipy_b200.plan as hp
orig_scatter = hp.dgemm_scatter
def timed_scatter(*a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); orig_scatter(*a, **k); e1.record(); marks.append(("scatter-gemm M=%d N=%d K=%d b=%d" % (a[0], a[1], a[2], k.get("batch", 1)), e0, e1))
hp.dgemm_scatter = timed_scatter
for it in range(4):
    marks.clear()
    dist.barrier(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
    t0.record(); c = T.forward(f); t1.record(); gout = T.backward(c); t2.record()
    torch.cuda.synchronize()
if rank == 0:
    print("world %d mode %s: forward %.1f us, backward %.1f us" % (world, T.exchange, t0.elapsed_time(t1) * 1e3, t1.elapsed_time(t2) * 1e3))
    for name, e0, e1 in marks:
        print("   %-50s %.1f us" % (name, e0.elapsed_time(e1) * 1e3))
dist.destroy_process_group()
