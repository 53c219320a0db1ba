"""Developer script (torchrun): where the time of the sharded parity-split 3-D round trip goes -- graph-replayed round
trips of variants that are NOT all correct (timing only): scatter-GEMM tile configurations, barriers removed, stores
kept local.    torchrun --nproc-per-node N tools/dist_variants.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import gauss_hermite
from hermipy_b200.plan import Plan
from hermipy_b200.dist import ShardedParityTransform, plan_parity_tables, DeviceOps

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256
x, w = gauss_hermite(n)
plan = Plan(n - 1, [x] * 3, [w] * 3, index_set="cube")
tables = plan_parity_tables(plan)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")


class CfgOps(DeviceOps):
    def __init__(self, scatter_cfg=None, plain_cfg=None):
        self.scatter_cfg, self.plain_cfg = scatter_cfg, plain_cfg

    def gemm(self, g):
        cfg = self.scatter_cfg if g.blocks else self.plain_cfg
        if cfg is not None:
            os.environ["HB_GEMM_CFG"] = str(cfg)
        try:
            super().gemm(g)
        finally:
            os.environ.pop("HB_GEMM_CFG", None)


class BarrierKernels(ShardedParityTransform):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.fused_flags = False


class FusedFlags(ShardedParityTransform):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.fused_flags = True


class NoSync(ShardedParityTransform):
    """neither flags nor barriers (timing only)"""
    def _exchange_flags(self, slot):
        return None, None

    def _before_scatter(self, name, hdl):
        pass


class LocalStores(BarrierKernels):
    def _peer_ptr(self, t, hdl, q):
        return t.data_ptr()


def run(label, cls=ShardedParityTransform, **kw):
    T = cls([n] * 3, [n] * 3, tables, exchange="p2p", ops=CfgOps(**kw))
    f = torch.randn(T.local_grid_shape(), dtype=torch.float64, device="cuda")
    graph, c, gout = T.capture(f)
    ts = []
    for _ in range(30):
        flush.zero_()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = torch.tensor([float(np.median(ts)), float(min(ts))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("world %d %-40s round trip %.1f us median, %.1f us min" % (world, label, t[0].item(), t[1].item()), flush=True)
    del graph, T
    torch.cuda.synchronize()
    dist.barrier()


run("baseline (bulk stores, barrier kernels)")
run("fused flags instead of barrier kernels", cls=FusedFlags)
os.environ["HB_GEMM_BULK"] = "0"
run("register epilogue (HB_GEMM_BULK=0)")
os.environ.pop("HB_GEMM_BULK")
run("scatter GEMMs bulk cfg 7", scatter_cfg=7)
run("no flags, no barriers (timing only)", cls=NoSync)
run("stores kept local (timing only)", cls=LocalStores)
dist.destroy_process_group()
