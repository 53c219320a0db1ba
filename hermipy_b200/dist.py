"""Multi-GPU composition of the hot path (one process per GPU, torch.distributed for the plumbing).

The reference is a single-threaded CPU library: nothing here has a reference counterpart except the
results.  Sharding follows the structure of each workload:

  * batched 1-D transforms   -- independent rows: shard the batch, no data-path collective;
  * d-dim (d >= 2) transform -- shard the OUTER grid axis; contract the inner axes locally, then ONE
    all-to-all re-shards from grid axis 0 to coefficient axis 1 so that axis 0 can be contracted
    locally; coefficients stay sharded (lexicographic box, split along coefficient axis 1);
  * varf                     -- independent output rows: each rank assembles its own row block
    (tiles that hold no owned row are skipped inside the GEMM), never gathered.

The host-side logic (partitioning, layouts, the exchange) also runs on CPU tensors with the gloo
backend so that it can be tested without GPUs: the tests inject a CPU stand-in for the GEMM (`gemm=`, defined
in tests/test_dist_gloo.py); the package itself ships no CPU arithmetic.
"""
import numpy as np
import torch
import torch.distributed as dist


def even_up(v):
    return (v + 1) & ~1


def split_range(total, parts, index):
    """Contiguous block `index` of `total` items split as evenly as possible over `parts`."""
    base, rem = divmod(total, parts)
    lo = index * base + min(index, rem)
    return lo, lo + base + (1 if index < rem else 0)


def device_gemm(A, B, out):
    """out[b] = A[b] @ B[b] through the DMMA/TMA kernel.  A: (M,K) or (b,M,K); B: (K,N) or (b,K,N);
    last dimension contiguous, leading dimensions even."""
    from .plan import dgemm
    batch = out.shape[0] if out.dim() == 3 else 1
    M, N = out.shape[-2], out.shape[-1]
    K = A.shape[-1]
    sA = A.stride(0) if A.dim() == 3 else 0
    sB = B.stride(0) if B.dim() == 3 else 0
    sC = out.stride(0) if out.dim() == 3 else 0
    dgemm(M, N, K, A, A.stride(-2), B, B.stride(-2), out, out.stride(-2), batch=batch, strideA=sA, strideB=sB,
          strideC=sC)


class ShardedTransform:
    """d-dimensional (d >= 3) transform with the outer grid axis sharded over the ranks of `group`.

    Grid functions: each rank holds f[x_0 in slab][x_1]...[x_{d-1}] (last axis padded to an even pitch).
    Coefficients: the lexicographic box c[m_0][m_1 in slice][m_2]...[m_{d-1}] (padded last axis), i.e.
    sharded along coefficient axis 1.  `tables[a]` = (phi[n_a, P_a], w_phi[n_a, P_a]) as torch tensors
    on the compute device (views of the plan's device tables, or CPU tensors for the gloo tests).

    exchange = "nccl": the axis-1 GEMMs write a send buffer laid out [dest][x0][m1 in slice][rest] and one
               all_to_all_single moves it;
    exchange = "p2p":  the axis-1 GEMM for destination r stores its tile rows straight into rank r's
               receive buffer through NVLink peer memory (torch symmetric memory gives the mapped
               pointers), so the all-to-all is the GEMM epilogue; two device-side barriers fence it.
    All buffers are allocated on first use and reused, so a step is a fixed launch sequence
    (capturable in a CUDA graph: see `capture`).
    """

    def __init__(self, n_pts, P, tables, group=None, gemm=device_gemm, exchange="nccl"):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_pts, self.P = list(n_pts), list(P)
        self.dim = len(n_pts)
        assert self.dim >= 3, "outer-axis sharding is implemented for dim >= 3 (2-D single transforms are replica-only)"
        assert self.n_pts[0] % self.world == 0, "outer grid axis must divide evenly over the ranks"
        self.slab = self.n_pts[0] // self.world
        self.Ps = -(-self.P[1] // self.world)  # coefficient-axis-1 slice width (last slices zero-padded)
        self.gemm = gemm
        self.np_last = even_up(self.n_pts[-1])
        self.Pp_last = even_up(self.P[-1])
        t = tables[0][0]
        self.dev, self.dtype = t.device, t.dtype
        self.exchange = exchange if self.world > 1 else "none"
        self._bufs = {}
        self._symm = {}
        # row-major copies of the tables with even pitch: phi[i][k], (w phi)^T[k][i], phi^T[k][i]
        self.phi = [self._padded(tables[a][0][:, :self.P[a]]) for a in range(self.dim)]
        self.wphi = [self._padded(tables[a][1][:, :self.P[a]]) for a in range(self.dim)]
        self.wphiT = [self._padded(tables[a][1][:, :self.P[a]].t()) for a in range(self.dim)]
        self.phiT = [self._padded(tables[a][0][:, :self.P[a]].t()) for a in range(self.dim)]
        self._graphs = {}

    # -- buffers ----------------------------------------------------------------------------------
    def _padded(self, t):
        out = torch.zeros((t.shape[0], even_up(t.shape[1])), dtype=self.dtype, device=self.dev)
        out[:, :t.shape[1]] = t
        return out[:, :t.shape[1]]

    def _buf(self, name, *shape):
        key = (name, shape)
        if key not in self._bufs:
            self._bufs[key] = torch.zeros(shape, dtype=self.dtype, device=self.dev)
        return self._bufs[key]

    def _symm_buf(self, name, *shape):
        """symmetric-memory receive buffer + rendezvous handle (p2p exchange)"""
        key = (name, shape)
        if key not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(*shape, dtype=self.dtype, device=self.dev)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            self._symm[key] = (t, hdl)
        return self._symm[key]

    def local_grid_shape(self):
        return (self.slab, *self.n_pts[1:-1], self.np_last)

    def local_coef_shape(self):
        return (self.P[0], self.Ps, *self.P[2:-1], self.Pp_last)

    # -- the exchange (grid axis 0 <-> coefficient axis 1) -------------------------------------------
    def _peer_ptrs(self, hdl, byte_offset):
        return [int(hdl.buffer_ptrs[r]) + byte_offset for r in range(self.world)]

    @staticmethod
    def _clip(block, A):
        """rows of the destination block actually produced by A (the last slice may be short)"""
        rows = A.shape[-2]
        return block[..., :rows, :] if block.shape[-2] != rows else block

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, f):
        d, g, s, Ps = self.dim, self.world, self.slab, self.Ps
        cur = f.reshape(s, *self.n_pts[1:-1], self.np_last)
        for a in range(d - 1, 1, -1):            # local contractions of axes d-1 ... 2
            cur = self._contract(cur, a, forward=True)
        rest_shape = tuple(cur.shape[2:])
        rest = int(np.prod(rest_shape))
        cur2 = cur.reshape(s, self.n_pts[1], rest)
        # axis 1: G[x0][m1][rest] = (w phi_1)^T @ cur[x0], delivered as [src][x0l][m1 in slice][rest]
        if self.exchange == "p2p":
            from .plan import dgemm_scatter
            recv, hdl = self._symm_buf("fwd", g, s, Ps, rest)
            hdl.barrier(channel=0)               # every rank has finished reading its previous contents
            blocks = self._peer_ptrs(hdl, self.rank * s * Ps * rest * 8)
            A = self.wphiT[1]
            dgemm_scatter(self.P[1], rest, self.n_pts[1], A, A.stride(0), cur2, rest, blocks, Ps, rest, batch=s,
                          strideA=0, strideB=cur2.stride(0), strideC=Ps * rest)
            hdl.barrier(channel=1)               # every peer's stores have landed
        else:
            send = self._buf("fwd_send", g, s, Ps, rest)
            for r in range(g):
                lo, hi = r * Ps, min((r + 1) * Ps, self.P[1])
                if hi > lo:
                    self.gemm(self.wphiT[1][lo:hi], cur2, send[r, :, :hi - lo])
            if self.exchange == "none":
                recv = send
            else:
                recv = self._buf("fwd_recv", g, s, Ps, rest)
                dist.all_to_all_single(recv, send, group=self.group)
        cols = Ps * rest
        out = self._buf("fwd_out", self.P[0], cols)
        self.gemm(self.wphiT[0], recv.reshape(self.n_pts[0], cols), out)
        return out.reshape(self.P[0], Ps, *rest_shape)

    def _contract(self, cur, a, forward):
        """contract axis a (>= 2) of cur[x0][x1]...[axis a][cols]"""
        d = self.dim
        n_in, n_out = (self.n_pts[a], self.P[a]) if forward else (self.P[a], self.n_pts[a])
        tag = "f" if forward else "b"
        if a == d - 1:
            p_in, p_out = (self.np_last, self.Pp_last) if forward else (self.Pp_last, self.np_last)
            rows = int(np.prod(cur.shape[:-1]))
            out = self._buf("c%s%d" % (tag, a), rows, p_out)
            table = self.wphi[a] if forward else self.phiT[a]          # [n_in][n_out]
            self.gemm(cur.reshape(rows, p_in)[:, :n_in], table, out[:, :n_out])
            return out.reshape(*cur.shape[:-1], p_out)
        lead = int(np.prod(cur.shape[:a]))
        cols = int(np.prod(cur.shape[a + 1:]))
        out = self._buf("c%s%d" % (tag, a), lead, n_out, cols)
        table = self.wphiT[a] if forward else self.phi[a]              # [n_out][n_in]
        self.gemm(table, cur.reshape(lead, n_in, cols), out)
        return out.reshape(*cur.shape[:a], n_out, *cur.shape[a + 1:])

    # -- backward ---------------------------------------------------------------------------------
    def backward(self, c):
        d, g, s, Ps = self.dim, self.world, self.slab, self.Ps
        rest_shape = tuple(c.shape[2:])
        rest = int(np.prod(rest_shape))
        cols = Ps * rest
        c2 = c.reshape(self.P[0], cols)
        # axis 0: X[x0][(ml, rest)] = phi_0 @ c; the rows of rank r's slab are delivered to rank r, next to
        # the slices of the other ranks: Y[x0l][m1 = (src, ml)][rest]
        if self.exchange == "p2p":
            from .plan import dgemm_scatter
            Y, hdl = self._symm_buf("bwd", s, g * Ps, rest)
            hdl.barrier(channel=0)
            blocks = self._peer_ptrs(hdl, self.rank * cols * 8)
            A = self.phi[0]
            dgemm_scatter(self.n_pts[0], cols, self.P[0], A, A.stride(0), c2, c2.stride(0), blocks, s, g * cols)
            hdl.barrier(channel=1)
        else:
            send = self._buf("bwd_send", g, s, cols)
            for r in range(g):
                self.gemm(self.phi[0][r * s:(r + 1) * s], c2, send[r])
            if self.exchange == "none":
                recv = send
            else:
                recv = self._buf("bwd_recv", g, s, cols)
                dist.all_to_all_single(recv, send, group=self.group)
            Y = self._buf("bwd_Y", s, g * Ps, rest)
            Y.copy_(recv.reshape(g, s, Ps, rest).permute(1, 0, 2, 3).reshape(s, g * Ps, rest))
        cur = self._buf("bwd_x1", s, self.n_pts[1], rest)
        self.gemm(self.phi[1], Y[:, :self.P[1]], cur)                    # [x0l][x1][rest]
        cur = cur.reshape(s, self.n_pts[1], *rest_shape)
        for a in range(2, d):
            cur = self._contract(cur, a, forward=False)
        return cur

    # -- CUDA graph of a forward+backward round trip on fixed buffers ---------------------------------
    def capture(self, f):
        """Warm up (allocates every buffer), then capture forward(f) -> backward as one CUDA graph.
        Returns (graph, coefficient tensor, grid tensor); graph.replay() re-runs the launch sequence."""
        for _ in range(2):
            c = self.forward(f)
            gout = self.backward(c)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(self.group)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            c = self.forward(f)
            gout = self.backward(c)
        return graph, c, gout


def plan_tables_as_tensors(plan):
    """Views of a Plan's per-axis device tables as torch CUDA tensors [(phi, w_phi)] (no copy)."""
    out = []
    for a in range(plan.dim):
        pair = []
        for kind in (0, 1):
            ptr, rows, cols, ld = plan.table(a, kind)
            pair.append(_tensor_from_ptr(ptr, rows, ld)[:, :cols])
        out.append(tuple(pair))
    return out


def _tensor_from_ptr(ptr, rows, ld):
    """Wrap device memory owned by the plan as a (rows, ld) float64 CUDA tensor via __cuda_array_interface__."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (rows, ld), "typestr": "<f8", "data": (int(ptr), False), "version": 2,
                                  "strides": None}
    return torch.as_tensor(h, device="cuda")
