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
import os

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
    def _peer_ptrs(self, hdl, byte_offset, t=None):
        # a handle may describe a larger allocation that several tensors were carved from: offset of `t` in it
        base = 0 if t is None else t.data_ptr() - int(hdl.buffer_ptrs[self.rank])
        return [int(hdl.buffer_ptrs[r]) + base + byte_offset for r in range(self.world)]

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
            blocks = self._peer_ptrs(hdl, self.rank * s * Ps * rest * 8, recv)
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
            blocks = self._peer_ptrs(hdl, self.rank * cols * 8, Y)
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


# ------------------------------------------------------------------------------------------------------
# Parity-split sharded transform: the production path for symmetric (Gauss-Hermite) grids.
class Gemm:
    """One launch of the two-level-batch GEMM (hb_gemm_desc of include/hermite_b200.h) on torch tensors:
    operand X is the flat tensor `X` from element offset `x_off` on.  `blocks` = [(tensor or device address,
    element offset)] turns the store into the row-block scatter."""

    def __init__(self, M, N, K, A, lda, B, ldb, C=None, ldc=0, a_off=0, b_off=0, c_off=0, batch=1, batch_lo=1,
                 strideA=0, strideA_lo=0, strideB=0, strideB_lo=0, strideC=0, strideC_lo=0, blocks=None, block_rows=0,
                 row_offset_lo=0, signal=None, wait=None):
        # signal = ([peer flag addresses], local state address): the last CTA publishes "stores landed" to the peers;
        # wait = (local flags address, count, local epoch address): the kernel waits for the peers' flags first
        self.__dict__.update(locals())
        del self.__dict__["self"]


class DeviceOps:
    """The CUDA implementation of the three primitives the sharded transform is composed of."""

    def gemm(self, g):
        import ctypes as C
        from ._lib import GemmDesc, check, lib
        from .plan import _stream
        ptr = lambda t, off: (t.data_ptr() if hasattr(t, "data_ptr") else int(t)) + 8 * off
        d = GemmDesc()
        d.M, d.N, d.K, d.batch, d.batch_lo, d.a_mmajor = g.M, g.N, g.K, g.batch, g.batch_lo, 0
        d.A, d.lda, d.strideA, d.strideA_lo = ptr(g.A, g.a_off), g.lda, g.strideA, g.strideA_lo
        d.B, d.ldb, d.strideB, d.strideB_lo = ptr(g.B, g.b_off), g.ldb, g.strideB, g.strideB_lo
        d.ldc, d.strideC, d.strideC_lo = g.ldc, g.strideC, g.strideC_lo
        keep = None
        if g.blocks:
            keep = (C.c_void_p * len(g.blocks))(*[ptr(t, off) for t, off in g.blocks])
            d.blocks = C.cast(keep, C.POINTER(C.c_void_p))
            d.n_blocks, d.block_rows, d.row_offset, d.row_offset_lo = len(g.blocks), g.block_rows, 0, g.row_offset_lo
            d.C = None
        else:
            d.C = ptr(g.C, g.c_off)
        keep2 = None
        if g.signal is not None:
            keep2 = (C.c_void_p * len(g.signal[0]))(*g.signal[0])
            d.signal_flags, d.n_signal, d.signal_state = C.cast(keep2, C.POINTER(C.c_void_p)), len(g.signal[0]), g.signal[1]
        if g.wait is not None:
            d.wait_flags, d.n_wait, d.wait_epoch = g.wait
        check(lib().hb_dev_dgemm_ex(C.byref(d), _stream()))

    def parity_pass(self, merge, n, split, nat_stride, srt_stride, srt_par, src, dst, nat_pad_last=0):
        import ctypes as C
        from ._lib import c_ip, c_llp, check, lib
        from .plan import _stream
        ia = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        la = lambda v: np.ascontiguousarray(v, dtype=np.int64)
        n_, sp_, ns_, ss_, sq_ = ia(n), ia(split), la(nat_stride), la(srt_stride), la(srt_par)
        check(lib().hb_dev_parity_pass(int(merge), len(n), n_.ctypes.data_as(c_ip), sp_.ctypes.data_as(c_ip),
                                       ns_.ctypes.data_as(c_llp), ss_.ctypes.data_as(c_llp), sq_.ctypes.data_as(c_llp),
                                       1, 0, 0, int(nat_pad_last), src.data_ptr(), dst.data_ptr(), _stream()))


def parity_tables(phi, wphi):
    """Per-parity tables of one symmetric axis from phi[i][k], (w phi)[i][k] (layout only, no arithmetic):
    cW / cPhi [2][nh][Pep] and cWT / cPhiT [2][Pe][nhp], odd entry zero-padded, as hb_plan_parity_table lays them
    out.  Used by the CPU (gloo) tests; on the GPU the plan supplies the same tables (plan_parity_tables)."""
    n, P = phi.shape
    nh, Pe = (n + 1) // 2, (P + 1) // 2
    nhp, Pep, lo = even_up(nh), even_up(Pe), n - nh
    out = {}
    for name, t in (("cW", wphi), ("cPhi", phi)):
        a = torch.zeros((2, nh, Pep), dtype=t.dtype, device=t.device)
        b = torch.zeros((2, Pe, nhp), dtype=t.dtype, device=t.device)
        for par in range(2):
            sub = t[lo:, par::2]
            a[par, :, :sub.shape[1]] = sub
            b[par, :sub.shape[1], :nh] = sub.t()
        out[name], out[name + "T"] = a, b
    return out


def plan_parity_tables(plan):
    """[{cW, cWT, cPhi, cPhiT}] per axis as torch views of the plan's device tables (no copy)."""
    import ctypes as C
    from ._lib import check, lib
    from .plan import _stream
    res = []
    for a in range(plan.dim):
        tabs = {}
        for kind, name in enumerate(("cW", "cWT", "cPhi", "cPhiT")):
            ptr, r, c, ld, st = C.c_void_p(), C.c_longlong(0), C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
            check(lib().hb_plan_parity_table(plan._h, a, kind, C.byref(ptr), C.byref(r), C.byref(c), C.byref(ld),
                                             C.byref(st), _stream()))
            tabs[name] = _tensor_from_ptr(ptr.value, 2 * r.value, ld.value).reshape(2, r.value, ld.value)
        res.append(tabs)
    return res


class ShardedParityTransform:
    """3-D transform on a symmetric (Gauss-Hermite) grid with the OUTER grid axis sharded over the ranks of
    `group`, every axis contraction split by parity (half the flops), one exchange per direction.

    Ownership.  The parity split pairs node x with -x, so rank r owns a SYMMETRIC pair of half-slabs of axis 0:
    with nh0 = n0/2 and sh = nh0/g, the nodes n0/2 - (r+1) sh ... n0/2 - r sh - 1 and their mirror images
    n0/2 + r sh ... n0/2 + (r+1) sh - 1 (`local_x0()` lists them; local array [2 sh][n1][n2 padded to even]).
    Index sets other than the cube are handled through the box: `gather_coefficients` / `scatter_coefficients` move
    between a coefficient vector in the reference's enumeration of ANY set contained in the box (triangle, cross, ...)
    and the sharded box, whose cells outside the set are computed and dropped / zero.
    Coefficients stay sharded: the parity-sorted box c[m0'][m1' in slice][m2'] where an index k' of a sorted axis
    runs over the even basis functions first, then the odd ones, both blocks padded to the same size (`coef_index()`
    maps the local block to basis indices; pad cells are -1 and hold zeros); the slice of rank r is rows
    r Ps ... (r+1) Ps - 1 of the sorted axis 1.

    Launch sequence per direction: one parity pass over the local grid values (all three axes at once), three GEMM
    launches (even and odd halves are the two low-level batch entries of one launch), and the exchange, which is
    the EPILOGUE of the GEMM that precedes it: that GEMM stores each row block straight into the destination
    rank's buffer (NVLink peer memory through torch symmetric memory, `exchange="p2p"`, fenced by ONE device
    barrier), or into a send buffer moved by one all_to_all_single (`exchange="nccl"`, also what the gloo tests
    run).  All buffers are allocated once, so a round trip is a fixed launch sequence (see `capture`).
    """

    def __init__(self, n_pts, P, tables, group=None, ops=None, exchange="nccl", emulate=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if emulate is not None:     # developer switch: the launch sequence of rank 0 of `emulate` ranks on ONE GPU, the
            self.world, self.rank, exchange = int(emulate), 0, "none"   # exchange kept local (timing only, results wrong)
        assert len(n_pts) == 3 and len(P) == 3, "the parity-split sharded transform is 3-D"
        self.n, self.P = [int(v) for v in n_pts], [int(v) for v in P]
        g = self.world
        assert self.n[0] % 2 == 0 and (self.n[0] // 2) % g == 0, \
            "outer axis: an even number of nodes whose half divides evenly over the ranks"
        self.ops = ops if ops is not None else DeviceOps()
        self.tables = tables
        t = tables[0]["cW"]
        self.dev, self.dtype = t.device, t.dtype
        self.exchange = exchange if (emulate is None and (g > 1 or (exchange == "p2p" and dist.is_initialized()))) else "none"
        n, Pn = self.n, self.P
        self.nh = [(v + 1) // 2 for v in n]
        self.Pe = [(v + 1) // 2 for v in Pn]
        self.nhp = [even_up(v) for v in self.nh]
        self.Pep = [even_up(v) for v in self.Pe]
        self.sh = self.nh[0] // g
        self.n2p = even_up(n[2])
        self.G1, self.G2 = 2 * self.nh[1], 2 * self.nhp[2]
        self.C0, self.C1, self.C2 = 2 * self.Pe[0], 2 * self.Pe[1], 2 * self.Pep[2]
        self.Ps = -(-self.C1 // g)
        self._bufs, self._symm, self._pending = {}, {}, set()
        self.always_fence = bool(int(os.environ.get("HB_DIST_ALWAYS_FENCE", "0")))   # developer switch
        # p2p: the exchange is fenced by ONE symmetric-memory barrier kernel between the scatter GEMM and its consumer.
        # HB_DIST_FUSED_FLAGS=1 replaces it by completion flags written by the scatter GEMM's last CTA and awaited by
        # the consumer GEMM's first operand load (no kernel in between); measured on 2 GPUs it is the slower of the
        # two (561 vs 550 us per round trip, 542 with no fence at all), so it is off by default.
        self.fused_flags = bool(int(os.environ.get("HB_DIST_FUSED_FLAGS", "0")))

    # -- layout -----------------------------------------------------------------------------------
    def local_x0(self):
        """global indices of the axis-0 nodes this rank owns, in local order"""
        half, sh, r = self.n[0] // 2, self.sh, self.rank
        return np.concatenate([np.arange(half - (r + 1) * sh, half - r * sh), np.arange(half + r * sh, half + (r + 1) * sh)])

    def local_grid_shape(self):
        return (2 * self.sh, self.n[1], self.n2p)

    def local_coef_shape(self):
        return (self.C0, self.Ps, self.C2)

    def coef_index(self):
        """(m0, m1, m2): basis index of every row / local slice row / column of the local coefficient block (-1 = pad)"""
        def sorted_axis(P, blk):
            out = np.full(2 * blk, -1, dtype=np.int64)
            out[:(P + 1) // 2] = np.arange(0, P, 2)
            out[blk:blk + P // 2] = np.arange(1, P, 2)
            return out
        m1 = np.full(self.world * self.Ps, -1, dtype=np.int64)
        m1[:self.C1] = sorted_axis(self.P[1], self.Pe[1])
        return (sorted_axis(self.P[0], self.Pe[0]), m1[self.rank * self.Ps:(self.rank + 1) * self.Ps],
                sorted_axis(self.P[2], self.Pep[2]))

    def _buf(self, name, size):
        if name not in self._bufs:
            self._bufs[name] = torch.zeros(size, dtype=self.dtype, device=self.dev)
        return self._bufs[name]

    def _symm_buf(self, name, size):
        if name not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(size, dtype=self.dtype, device=self.dev)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            hdl.barrier(channel=0)     # no peer may store into this buffer before its owner's zero-fill has run
            self._symm[name] = (t, hdl)
        return self._symm[name]

    def _exchange_flags(self, slot):
        """(signal, wait) arguments of the scatter GEMM / its consumer for exchange `slot` (0 forward, 1 backward):
        flags[slot][src] lives in symmetric memory on every rank, the epoch / CTA counters in local memory."""
        g, r = self.world, self.rank
        if "flags" not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem
            t = symm_mem.empty(2 * g, dtype=torch.int64, device=self.dev)
            t.zero_()
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            hdl.barrier(channel=0)
            self._symm["flags"] = (t, hdl)
            self._bufs["flag_state"] = torch.zeros(4, dtype=torch.int64, device=self.dev)
        flags, hdl = self._symm["flags"]
        state = self._bufs["flag_state"]
        signal = ([self._peer_ptr(flags, hdl, q) + 8 * (slot * g + r) for q in range(g)], state.data_ptr() + 16 * slot)
        wait = (flags.data_ptr() + 8 * slot * g, g, state.data_ptr() + 16 * slot)
        return signal, wait

    def _peer_ptr(self, t, hdl, q):
        """address of rank q's copy of the symmetric tensor `t` (a handle may describe a larger allocation that
        several tensors were carved from: `buffer_ptrs` are then the block bases, not the tensors)"""
        return int(hdl.buffer_ptrs[q]) + (t.data_ptr() - int(hdl.buffer_ptrs[self.rank]))

    def _parity_args(self):
        sh, n, S1 = self.sh, self.n, self.G1 * self.G2
        return dict(n=[2 * sh, n[1], n[2]], split=[1, 1, 1], nat_stride=[n[1] * self.n2p, self.n2p, 1],
                    # axis 0 interleaves parity and h ([h0][p0]) so that (h0, p0) is ONE uniformly strided batch index
                    srt_stride=[2 * S1, self.G2, 1], srt_par=[S1, self.nh[1] * self.G2, self.nhp[2]])

    # -- exchange fences (p2p) -----------------------------------------------------------------------
    # A peer buffer may be overwritten once every rank has finished reading its previous contents.  Any barrier
    # passed since those reads guarantees that (ranks run the same launch sequence in stream order), so a round
    # trip needs ONE barrier per direction -- the one that publishes the stores -- and a second one only when the
    # same direction runs twice in a row.
    def _before_scatter(self, name, hdl):
        if name in self._pending or self.always_fence:
            hdl.barrier(channel=0)
            self._pending.clear()

    def _after_scatter(self, name, hdl):
        if not self.fused_flags:
            hdl.barrier(channel=1)
        self._pending.clear()
        self._pending.add(name)     # about to be read locally

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, f):
        g, r, sh, Ps, ops, T = self.world, self.rank, self.sh, self.Ps, self.ops, self.tables
        nh, nhp, Pe, Pep, G1, G2, C2 = self.nh, self.nhp, self.Pe, self.Pep, self.G1, self.G2, self.C2
        X0 = 2 * sh
        fs = self._buf("fs", X0 * G1 * G2)
        ops.parity_pass(0, src=f, dst=fs, **self._parity_args())
        # axis 2: t1[(X0, X1)][p2 Pep2 + m'] = sum_h fs[(X0, X1)][p2 nhp2 + h] cW2[p2][h][m']
        t1 = self._buf("t1", X0 * G1 * C2)
        ops.gemm(Gemm(X0 * G1, Pe[2], nh[2], fs, G2, T[2]["cW"], Pep[2], t1, C2, batch_lo=2, strideA_lo=nhp[2],
                      strideB_lo=nh[2] * Pep[2], strideC_lo=Pep[2]))
        # axis 1 + exchange: sorted row m1' = p1 Pe1 + j of batch entry X0 = (h0l, p0) goes to rank m1' // Ps, which
        # receives R[h0 = src sh + h0l][p0][m1' % Ps][C2]
        blk = sh * 2 * Ps * C2
        signal = wait = None
        if self.exchange == "p2p":
            recv, hdl = self._symm_buf("F", g * blk)
            self._before_scatter("F", hdl)
            blocks = [(self._peer_ptr(recv, hdl, q), r * blk) for q in range(g)]
            if self.fused_flags:
                signal, wait = self._exchange_flags(0)
        else:
            send = self._buf("F_send", g * blk)
            blocks = [(send, q * blk) for q in range(g)]
        ops.gemm(Gemm(Pe[1], C2, nh[1], T[1]["cWT"], nhp[1], t1, C2, ldc=C2, batch=X0, batch_lo=2,
                      strideA_lo=Pe[1] * nhp[1], strideB=G1 * C2, strideB_lo=nh[1] * C2, strideC=Ps * C2,
                      blocks=blocks, block_rows=Ps, row_offset_lo=Pe[1], signal=signal))
        if self.exchange == "p2p":
            self._after_scatter("F", hdl)
        elif self.exchange == "nccl":
            recv = self._buf("F_recv", g * blk)
            dist.all_to_all_single(recv, send, group=self.group)
        else:
            recv = send
        # axis 0: c[p0 Pe0 + m'][Ps C2] = sum_h0 cWT0[p0][m'][h0] R[h0][p0][Ps C2]
        cols = Ps * C2
        c = self._buf("coef", self.C0 * cols)
        ops.gemm(Gemm(Pe[0], cols, nh[0], T[0]["cWT"], nhp[0], recv, 2 * cols, c, cols, batch_lo=2,
                      strideA_lo=Pe[0] * nhp[0], strideB_lo=cols, strideC_lo=Pe[0] * cols, wait=wait))
        return c.view(self.C0, Ps, C2)

    # -- backward ---------------------------------------------------------------------------------
    def backward(self, c):
        g, r, sh, Ps, ops, T = self.world, self.rank, self.sh, self.Ps, self.ops, self.tables
        nh, nhp, Pe, Pep, G1, G2, C2 = self.nh, self.nhp, self.Pe, self.Pep, self.G1, self.G2, self.C2
        X0, cols = 2 * sh, Ps * C2
        c = c.reshape(-1)
        # axis 0 + exchange: row h0 of parity p0 goes to rank h0 // sh, next to the slices of the other ranks:
        # Y[h0l][p0][m1' = src Ps + j][C2]
        signal = wait = None
        if self.exchange == "p2p":
            Y, hdl = self._symm_buf("B", X0 * g * cols)
            self._before_scatter("B", hdl)
            blocks = [(self._peer_ptr(Y, hdl, q), r * cols) for q in range(g)]
            ldc, lo_stride = 2 * g * cols, g * cols
            if self.fused_flags:
                signal, wait = self._exchange_flags(1)
        else:
            send = self._buf("B_send", g * X0 * cols)
            blocks = [(send, q * X0 * cols) for q in range(g)]
            ldc, lo_stride = 2 * cols, cols
        ops.gemm(Gemm(nh[0], cols, Pe[0], T[0]["cPhi"], Pep[0], c, cols, ldc=ldc, batch_lo=2,
                      strideA_lo=nh[0] * Pep[0], strideB_lo=Pe[0] * cols, strideC_lo=lo_stride, blocks=blocks,
                      block_rows=sh, signal=signal))
        if self.exchange == "p2p":
            self._after_scatter("B", hdl)
        else:
            if self.exchange == "nccl":
                recv = self._buf("B_recv", g * X0 * cols)
                dist.all_to_all_single(recv, send, group=self.group)
            else:
                recv = send
            Y = self._buf("B_Y", X0 * g * cols)
            Y.view(X0, g, cols).copy_(recv.view(g, X0, cols).permute(1, 0, 2))
        # axis 1: u[X0][p1 nh1 + h][C2] = sum_m' cPhi1[p1][h][m'] Y[X0][p1 Pe1 + m'][C2]
        u = self._buf("u", X0 * G1 * C2)
        ops.gemm(Gemm(nh[1], C2, Pe[1], T[1]["cPhi"], Pep[1], Y, C2, u, C2, batch=X0, batch_lo=2,
                      strideA_lo=nh[1] * Pep[1], strideB=g * cols, strideB_lo=Pe[1] * C2, strideC=G1 * C2,
                      strideC_lo=nh[1] * C2, wait=wait))
        # axis 2: v[(X0, X1)][p2 nhp2 + h] = sum_m' u[(X0, X1)][p2 Pep2 + m'] cPhiT2[p2][m'][h]
        v = self._buf("v", X0 * G1 * G2)
        ops.gemm(Gemm(X0 * G1, nh[2], Pe[2], u, C2, T[2]["cPhiT"], nhp[2], v, G2, batch_lo=2, strideA_lo=Pep[2],
                      strideB_lo=Pe[2] * nhp[2], strideC_lo=nhp[2]))
        out = self._buf("grid", X0 * self.n[1] * self.n2p)
        ops.parity_pass(1, src=v, dst=out, nat_pad_last=self.n2p, **self._parity_args())
        return out.view(X0, self.n[1], self.n2p)

    def capture(self, f):
        """Warm up (allocates every buffer), then capture forward(f) -> backward as one CUDA graph.
        Returns (graph, coefficient tensor, grid tensor); graph.replay() re-runs the launch sequence."""
        for _ in range(2):
            c = self.forward(f)
            gout = self.backward(c)
        torch.cuda.synchronize()
        if self.world > 1 and dist.is_initialized():
            dist.barrier(self.group)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            c = self.forward(f)
            gout = self.backward(c)
        return graph, c, gout

    # -- host buffers (the end-to-end path: this rank's slab in, this rank's coefficient block out) --------
    def forward_host(self, f_host, c_host):
        """f_host: page-locked CPU tensor of local_grid_shape(); c_host: page-locked CPU tensor of local_coef_shape().
        Upload, transform, download; returns after the coefficients have landed in c_host."""
        dev_f = self._buf("host_f", int(np.prod(self.local_grid_shape()))).view(self.local_grid_shape())
        dev_f.copy_(f_host, non_blocking=True)
        c_host.copy_(self.forward(dev_f), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return c_host

    def backward_host(self, c_host, f_host):
        dev_c = self._buf("host_c", int(np.prod(self.local_coef_shape()))).view(self.local_coef_shape())
        dev_c.copy_(c_host, non_blocking=True)
        f_host.copy_(self.backward(dev_c), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return f_host

    # -- the reference's enumeration -------------------------------------------------------------------
    def set_positions(self, degree, index_set="cube"):
        """(valid, pos): for every cell of the local coefficient block whether it holds a basis function of the
        index set, and the position of that function in the reference's enumeration (iterators.cpp order)."""
        key = ("pos", degree, index_set)
        if key not in self._bufs:
            from . import core
            m0, m1, m2 = self.coef_index()
            M0, M1, M2 = np.meshgrid(m0, m1, m2, indexing="ij")
            inside = (M0 >= 0) & (M1 >= 0) & (M2 >= 0)
            pos = np.full(inside.shape, -1, dtype=np.int64)
            pos[inside] = core.iterator_positions(np.stack([M0[inside], M1[inside], M2[inside]], axis=1), degree, index_set)
            valid = pos >= 0
            self._bufs[key] = (torch.from_numpy(valid).to(self.dev), torch.from_numpy(pos[valid]).to(self.dev))
        return self._bufs[key]

    def gather_coefficients(self, c, degree, index_set="cube", n_polys=None):
        """The sharded coefficient blocks as ONE vector in the reference's enumeration order, on every rank: each rank
        places its block (disjoint positions) and one all-reduce over NVLink assembles the vector -- the only
        collective of the transform path besides the exchange, and only for callers that want the whole vector."""
        valid, pos = self.set_positions(degree, index_set)
        if n_polys is None:
            from . import core
            n_polys = core.iterator_size(3, degree, index_set)
        full = torch.zeros(n_polys, dtype=self.dtype, device=self.dev)
        full[pos] = c[valid]
        if self.world > 1:
            dist.all_reduce(full, group=self.group)
        return full

    def scatter_coefficients(self, full, degree, index_set="cube"):
        """Inverse of `gather_coefficients`: this rank's block of the parity-sorted coefficient box from a coefficient
        vector in the reference's enumeration order (any index set contained in the box: cells outside the set are
        zero), ready for `backward`.  No communication: every rank picks its own cells."""
        valid, pos = self.set_positions(degree, index_set)
        c = torch.zeros(self.local_coef_shape(), dtype=self.dtype, device=self.dev)
        c[valid] = full[pos]
        return c


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
