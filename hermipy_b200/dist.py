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
backend so that it can be tested without GPUs: pass `gemm=reference_gemm`.
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


def reference_gemm(A, B, out):
    """CPU stand-in with the device GEMM's semantics (out[b] = A[b] @ B[b]); host-logic tests only."""
    torch.matmul(A, B, out=out)


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
    """d-dimensional (d >= 2) transform with the outer grid axis sharded over the ranks of `group`.

    Grid functions: each rank holds f[x_0 in slab][x_1]...[x_{d-1}] (last axis padded to an even pitch).
    Coefficients: the lexicographic box c[m_0][m_1 in slice][m_2]...[m_{d-1}] (padded last axis), i.e.
    sharded along coefficient axis 1.  `tables[a]` = (phi[n_a, P_a], w_phi[n_a, P_a]) as torch tensors
    on the compute device with even row pitch (views of the plan's device tables, or CPU tensors).
    """

    def __init__(self, n_pts, P, tables, group=None, gemm=device_gemm):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_pts, self.P = list(n_pts), list(P)
        self.dim = len(n_pts)
        assert self.dim >= 3, "outer-axis sharding is implemented for dim >= 3 (2-D single transforms are replica-only)"
        assert self.n_pts[0] % self.world == 0, "outer grid axis must divide evenly over the ranks"
        self.slab = self.n_pts[0] // self.world
        self.Ps = -(-self.P[1] // self.world)  # coefficient-axis-1 slice width (last slices zero-padded)
        self.tables = tables
        self.gemm = gemm
        self.np_last = even_up(self.n_pts[-1])
        self.Pp_last = even_up(self.P[-1])
        t = tables[0][0]
        self.dev, self.dtype = t.device, t.dtype

    # -- helpers ----------------------------------------------------------------------------------
    def _empty(self, *shape):
        return torch.zeros(shape, dtype=self.dtype, device=self.dev)

    def _exchange(self, send):
        if self.world == 1:
            return send
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def local_grid_shape(self):
        return (self.slab, *self.n_pts[1:-1], self.np_last)

    def local_coef_shape(self):
        return (self.P[0], self.Ps, *self.P[2:-1], self.Pp_last)

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, f):
        d, g, s, Ps = self.dim, self.world, self.slab, self.Ps
        cur = f.reshape(s, *self.n_pts[1:-1], self.np_last)
        # local contractions: axes d-1 ... 2 (if any)
        for a in range(d - 1, 1, -1):
            cur = self._contract_forward(cur, a)
        # axis 1, written straight into the send layout [dest][x0][m1 in slice][rest]
        rest = int(np.prod(cur.shape[2:]))
        cur2 = cur.reshape(s, self.n_pts[1], rest)
        send = self._empty(g, s, Ps, rest)
        wT = self._rowmajor(self._wphi(1).t())                                          # [P1][n1]
        for r in range(g):
            lo, hi = r * Ps, min((r + 1) * Ps, self.P[1])
            if hi > lo:
                self.gemm(wT[lo:hi], cur2, send[r, :, :hi - lo])
        recv = self._exchange(send)                                                    # [src][x0l][ml][rest]
        cols = Ps * rest
        cols_p = even_up(cols)
        R = recv.reshape(self.n_pts[0], cols)
        if cols_p != cols:
            Rp = self._empty(self.n_pts[0], cols_p)
            Rp[:, :cols] = R
            R = Rp
        out = self._empty(self.P[0], cols_p)
        self.gemm(self._rowmajor(self._wphi(0).t()), R, out)
        return out[:, :cols].reshape(self.P[0], Ps, *cur.shape[2:])

    def _wphi(self, a):
        return self.tables[a][1][:, :self.P[a]]

    def _phi(self, a):
        return self.tables[a][0][:, :self.P[a]]

    def _rowmajor(self, t):
        """contiguous copy with an even row pitch (tables are tiny; done once per call)"""
        out = self._empty(t.shape[0], even_up(t.shape[1]))
        out[:, :t.shape[1]] = t
        return out[:, :t.shape[1]]

    def _contract_forward(self, cur, a):
        """contract grid axis a (>= 2) of cur[x0][x1]...[x_a][cols] with w_phi_a -> [...][m_a][cols]"""
        d = self.dim
        if a == d - 1:
            rows = int(np.prod(cur.shape[:-1]))
            out = self._empty(rows, self.Pp_last)
            self.gemm(cur.reshape(rows, self.np_last)[:, :self.n_pts[a]], self._wphi(a), out[:, :self.P[a]])
            return out.reshape(*cur.shape[:-1], self.Pp_last)
        lead = int(np.prod(cur.shape[:a]))
        cols = int(np.prod(cur.shape[a + 1:]))
        out = self._empty(lead, self.P[a], cols)
        self.gemm(self._rowmajor(self._wphi(a).t()), cur.reshape(lead, self.n_pts[a], cols), out)
        return out.reshape(*cur.shape[:a], self.P[a], *cur.shape[a + 1:])

    # -- backward ---------------------------------------------------------------------------------
    def backward(self, c):
        d, g, s, Ps = self.dim, self.world, self.slab, self.Ps
        rest_shape = c.shape[2:]
        rest = int(np.prod(rest_shape))
        cols = Ps * rest
        X = self._empty(self.n_pts[0], cols)
        self.gemm(self._rowmajor(self._phi(0)), c.reshape(self.P[0], cols), X)          # [x0][ml][rest]
        recv = self._exchange(X.reshape(g, s, Ps, rest))                               # [src][x0l][ml][rest]
        Y = recv.permute(1, 0, 2, 3).reshape(s, g * Ps, rest)[:, :self.P[1]].contiguous()   # [x0l][m1][rest]
        cur = self._empty(s, self.n_pts[1], rest)
        self.gemm(self._rowmajor(self._phi(1)), Y, cur)                                 # [x0l][x1][rest]
        cur = cur.reshape(s, self.n_pts[1], *rest_shape)
        for a in range(2, d):
            cur = self._contract_backward(cur, a)
        return cur

    def _contract_backward(self, cur, a):
        d = self.dim
        if a == d - 1:
            rows = int(np.prod(cur.shape[:-1]))
            out = self._empty(rows, self.np_last)
            self.gemm(cur.reshape(rows, self.Pp_last)[:, :self.P[a]], self._rowmajor(self._phi(a).t()),
                      out[:, :self.n_pts[a]])
            return out.reshape(*cur.shape[:-1], self.np_last)
        lead = int(np.prod(cur.shape[:a]))
        cols = int(np.prod(cur.shape[a + 1:]))
        out = self._empty(lead, self.n_pts[a], cols)
        self.gemm(self._rowmajor(self._phi(a)), cur.reshape(lead, self.P[a], cols), out)
        return out.reshape(*cur.shape[:a], self.n_pts[a], *cur.shape[a + 1:])


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
