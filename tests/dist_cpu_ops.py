"""TEST INFRASTRUCTURE: CPU interpreters of the three device primitives the sharded transforms are composed of
(hb_dev_dgemm_ex, hb_dev_parity_pass), so that the host logic of hermipy_b200.dist -- partitioning, layouts, block
scatter, the exchange -- runs under gloo without a GPU.  They follow the semantics documented in
include/hermite_b200.h; the arithmetic itself is torch.matmul."""
import itertools

import numpy as np
import torch


class CpuOps:
    def gemm(self, g):
        for bh in range(g.batch):
            for bl in range(g.batch_lo):
                A = torch.as_strided(g.A, (g.M, g.K), (g.lda, 1), g.a_off + bh * g.strideA + bl * g.strideA_lo)
                B = torch.as_strided(g.B, (g.K, g.N), (g.ldb, 1), g.b_off + bh * g.strideB + bl * g.strideB_lo)
                R = A @ B
                if g.blocks:
                    for row in range(g.M):
                        d = bl * g.row_offset_lo + row
                        t, off = g.blocks[d // g.block_rows]
                        at = off + (d % g.block_rows) * g.ldc + bh * g.strideC + bl * g.strideC_lo
                        t[at:at + g.N] = R[row]
                else:
                    torch.as_strided(g.C, (g.M, g.N), (g.ldc, 1), g.c_off + bh * g.strideC + bl * g.strideC_lo).copy_(R)

    def parity_pass(self, merge, n, split, nat_stride, srt_stride, srt_par, src, dst, nat_pad_last=0):
        d = len(n)
        nh = [(v + 1) // 2 for v in n]
        src_np, dst_np = src.numpy(), dst.numpy()
        saxes = [a for a in range(d) if split[a]]
        for c in itertools.product(*[range(nh[a] if split[a] else n[a]) for a in range(d)]):
            nat0 = sum(c[a] * nat_stride[a] for a in range(d) if not split[a])
            srt0 = sum(c[a] * srt_stride[a] for a in range(d))
            pts = {}    # sign pattern -> natural offset (None if the mirror image coincides with the node itself)
            for signs in itertools.product((0, 1), repeat=len(saxes)):
                off, absent = nat0, False
                for s, a in zip(signs, saxes):
                    xp = n[a] - nh[a] + c[a]
                    xm = n[a] - 1 - xp
                    absent |= bool(s) and xm == xp
                    off += (xm if s else xp) * nat_stride[a]
                pts[signs] = None if absent else off
            for par in itertools.product((0, 1), repeat=len(saxes)):
                soff = srt0 + sum(p * srt_par[a] for p, a in zip(par, saxes))
                if not merge:
                    acc = 0.0
                    for signs, off in pts.items():
                        if off is not None:
                            acc += (-1) ** sum(p * s for p, s in zip(par, signs)) * src_np[off]
                    middle_odd = any(p and n[a] - 1 - 2 * (n[a] - nh[a] + c[a]) == 0 for p, a in zip(par, saxes))
                    dst_np[soff] = 0.0 * acc if middle_odd else acc
            if merge:
                for signs, off in pts.items():
                    if off is None:
                        continue
                    acc = 0.0
                    for par in itertools.product((0, 1), repeat=len(saxes)):
                        soff = srt0 + sum(p * srt_par[a] for p, a in zip(par, saxes))
                        acc += (-1) ** sum(p * s for p, s in zip(par, signs)) * src_np[soff]
                    dst_np[off] = acc
        if merge and nat_pad_last > n[-1]:
            rows = dst_np.reshape(-1, nat_pad_last)
            rows[:, n[-1]:] = 0.0
