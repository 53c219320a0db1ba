"""3-D Hermite transform with the outer grid axis sharded over the GPUs of one box (BASELINE configs[3] shape).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/sharded_transform_3d.py [n]

Every rank holds a symmetric pair of half-slabs of the grid function and ends with its block of the coefficients
(sharded along the parity-sorted coefficient axis 1).  The exchange between the two is the epilogue of a GEMM that
stores straight into the peers' buffers over NVLink.  `gather_coefficients` assembles the whole vector in the
reference's enumeration order for callers that want it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.dist import ShardedParityTransform, plan_parity_tables  # noqa: E402
from hermipy_b200.lib import hermegauss  # noqa: E402
from hermipy_b200.plan import Plan  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
degree = n - 1
x, w = hermegauss(n)
plan = Plan(degree, [x] * 3, [w] * 3, index_set="cube")            # per-axis tables (every rank builds its own)
T = ShardedParityTransform([n] * 3, [degree + 1] * 3, plan_parity_tables(plan), exchange="p2p" if world > 1 else "none")

x0 = T.local_x0()                                                   # the nodes of grid axis 0 this rank owns
f = torch.zeros(T.local_grid_shape(), dtype=torch.float64, device="cuda")
f[..., :n] = torch.from_numpy(np.exp(-(x[x0][:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2) / 8)
                              * (1 + 0.1 * x[None, :, None])).cuda()
c = T.forward(f)                                                    # local coefficient block (C0, Ps, C2)
g = T.backward(c)                                                   # back on this rank's slab
sw = torch.zeros_like(f)
sw[..., :n] = torch.from_numpy(np.sqrt(w[x0][:, None, None] * w[None, :, None] * w[None, None, :])).cuda()
sq = torch.stack([(((g - f) * sw) ** 2).sum(), ((f * sw) ** 2).sum()])
if world > 1:
    dist.all_reduce(sq)
full = T.gather_coefficients(c, degree, "cube")                     # whole vector, reference order, on every rank
assert torch.equal(T.scatter_coefficients(full, degree, "cube"), c)   # ... and back to this rank's block
tri = T.gather_coefficients(c, degree, "triangle")                 # any index set contained in the box
if rank == 0:
    print("%d GPU(s), %d^3 grid, degree %d: round trip error %.2e (weighted L2); |c_000| = %.6f; %d coefficients gathered "
          "(cube), %d (triangle)" % (world, n, degree, float((sq[0] / sq[1]).sqrt()), abs(float(full[0])), full.numel(), tri.numel()))
if world > 1:
    dist.destroy_process_group()
