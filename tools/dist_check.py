"""Developer script (torchrun): correctness of the sharded 3-D transform per exchange mode, with and
without CUDA graph capture, against the single-GPU plan result."""
import os, sys
import numpy as np, torch, torch.distributed as dist, scipy.special as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan
from hermipy_b200.dist import ShardedTransform, plan_tables_as_tensors

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = n
x, w = sp.roots_hermitenorm(n); w = w / np.sqrt(2 * np.pi)
plan = Plan(P - 1, [x] * 3, [w] * 3, index_set="cube")
g = np.exp(-(x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2) / 8) * (1 + 0.1 * x[None, :, None])
s = n // world
f = torch.from_numpy(np.ascontiguousarray(g[rank * s:(rank + 1) * s])).cuda()
ws = w[rank * s:(rank + 1) * s]
sw = torch.from_numpy(np.sqrt(ws[:, None, None] * w[None, :, None] * w[None, None, :])).cuda()
# single-GPU truth for the coefficients (lexicographic box)
F1 = plan.pack_grid(g[None]); C1 = plan.forward(F1)
lex = torch.zeros(P * P * P, dtype=torch.float64, device="cuda")
import ctypes
idx = torch.from_numpy(np.ascontiguousarray(__import__("hermipy_b200").core.iterator_list_indices(3, P - 1, "cube"))).cuda() if n <= 64 else None
tables = plan_tables_as_tensors(plan)
def gerr(gout):
    sq = torch.stack([(((gout.reshape(f.shape) - f) * sw) ** 2).sum(), ((f * sw) ** 2).sum()])
    dist.all_reduce(sq)
    return float((sq[0] / sq[1]).sqrt().item())
for mode in ("nccl", "p2p"):
    for use_graph in (False, True):
        T = ShardedTransform([n] * 3, [P] * 3, tables, exchange=mode)
        errs = []
        if use_graph:
            graph, c, gout = T.capture(f)
            for it in range(5):
                graph.replay(); torch.cuda.synchronize()
                errs.append(gerr(gout))
        else:
            for it in range(5):
                c = T.forward(f); gout = T.backward(c); torch.cuda.synchronize()
                errs.append(gerr(gout))
        cerr = -1.0
        if idx is not None:
            Ps = T.Ps
            mine = (idx[:, 1] >= rank * Ps) & (idx[:, 1] < (rank + 1) * Ps)
            sel = idx[mine]
            got = c[sel[:, 0], sel[:, 1] - rank * Ps, sel[:, 2]]
            want = C1[0, :plan.n_polys][mine]
            cerr = float(((got - want).norm() / want.norm()).item())
        t = torch.tensor([max(errs), cerr], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("mode=%s graph=%d: max round-trip err %.3e (per iter %s)  coef err %.3e" % (mode, use_graph, t[0].item(), ["%.1e" % e for e in errs], t[1].item()), flush=True)
dist.destroy_process_group()
