"""Developer script (python or torchrun): the parity-split sharded 3-D transform against the single-GPU plan --
coefficients of the local block, round trip, and the time of a graph-replayed round trip, per exchange mode.
    torchrun --nproc-per-node N tools/dist_parity_check.py [n]"""
import os, sys
import numpy as np, torch, torch.distributed as dist, scipy.special as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200 import core
from hermipy_b200.plan import Plan
from hermipy_b200.dist import ShardedParityTransform, plan_parity_tables

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = n
x, w = sp.roots_hermitenorm(n); w = w / np.sqrt(2 * np.pi)
x, w = (x - x[::-1]) / 2, (w + w[::-1]) / 2
plan = Plan(P - 1, [x] * 3, [w] * 3, index_set="cube")
g = np.exp(-(x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2) / 8) * (1 + 0.1 * x[None, :, None] + 0.05 * x[:, None, None] * x[None, None, :])
C1 = plan.forward(plan.pack_grid(g[None]))[0]          # single-GPU coefficients, set order
tables = plan_parity_tables(plan)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
for mode in (("nccl", "p2p") if world > 1 else ("none",)):
    T = ShardedParityTransform([n] * 3, [P] * 3, tables, exchange=mode)
    x0 = T.local_x0()
    f = torch.zeros(T.local_grid_shape(), dtype=torch.float64, device="cuda")
    f[..., :n] = torch.from_numpy(np.ascontiguousarray(g[x0])).cuda()
    sw = torch.zeros_like(f)
    sw[..., :n] = torch.from_numpy(np.sqrt(w[x0][:, None, None] * w[None, :, None] * w[None, None, :])).cuda()
    c = T.forward(f)
    m0, m1, m2 = T.coef_index()
    M0, M1, M2 = np.meshgrid(m0, m1, m2, indexing="ij")
    valid = (M0 >= 0) & (M1 >= 0) & (M2 >= 0)
    pos = core.iterator_positions(np.stack([M0[valid], M1[valid], M2[valid]], axis=1), P - 1, "cube")
    want = C1[torch.from_numpy(pos.astype(np.int64)).cuda()]
    got = c[torch.from_numpy(valid).cuda()]
    sq = torch.stack([((got - want) ** 2).sum(), (want ** 2).sum(), c[torch.from_numpy(~valid).cuda()].abs().sum()])
    if world > 1:
        dist.all_reduce(sq)
    cerr, padsum = float((sq[0] / sq[1]).sqrt()), float(sq[2])
    gout = T.backward(c)
    sq = torch.stack([(((gout - f) * sw) ** 2).sum(), ((f * sw) ** 2).sum()])
    if world > 1:
        dist.all_reduce(sq)
    rt = float((sq[0] / sq[1]).sqrt())
    graph, c, gout = T.capture(f)
    ts = []
    for _ in range(20):
        flush.zero_()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.median(ts)), float(min(ts))], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sq = torch.stack([(((gout - f) * sw) ** 2).sum(), ((f * sw) ** 2).sum()])
    if world > 1:
        dist.all_reduce(sq)
    if rank == 0:
        print("world %d mode %s n %d: coef rel err %.2e (pad sum %.1e), round trip %.2e, graph round trip %.2e; "
              "round trip %.1f us median, %.1f us min" % (world, mode, n, cerr, padsum, rt, float((sq[0] / sq[1]).sqrt()),
                                                         t[0].item() * 1e3, t[1].item() * 1e3), flush=True)
if world > 1:
    dist.destroy_process_group()
