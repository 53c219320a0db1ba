"""Developer script: a short, fixed launch sequence for `ncu --set full` captures (run via gpurun).

    python tools/profile_target.py c2        # 2 x (forward + inverse) C2 transform step, batch 17
    python tools/profile_target.py varf      # C2 cube varf: polynomial f (banded assembly) + dense f (Gram GEMM)
    python tools/profile_target.py c5        # one C5 forward+inverse over 8192 rows
    python tools/profile_target.py c4        # 2 x (forward + inverse) 256^3 transform
    python tools/profile_target.py tables    # recurrence kernel: basis tables of the C5 rule, psi table of the C3 rule
    python tools/profile_target.py matvec    # y = A x on a 40401 x 40401 resident matrix (the C2 cube varf size)
    python tools/profile_target.py gemm M N K CFG   # three launches of one plain GEMM with tile configuration CFG
"""
import os
import sys

import numpy as np
import scipy.special as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan  # noqa: E402


def gh(n):
    x, w = sp.roots_hermitenorm(n)
    return x, w / np.sqrt(2 * np.pi)


which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2":
    x, w = gh(401)
    pl = Plan(200, [x, x], [w, w], index_set="cube")
    F = torch.randn(17, pl.grid_size, dtype=torch.float64, device="cuda")
    C = pl.empty_coef(17)
    G = pl.empty_grid(17)
    for _ in range(2):
        pl.forward(F, C)
        pl.backward(C, G)
elif which == "varf":
    x, w = gh(401)
    X, Y = np.meshgrid(x, x, indexing="ij")
    V = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
    pl = Plan(200, [x, x], [w, w], index_set=sys.argv[2] if len(sys.argv) > 2 else "cube")
    out = torch.empty((pl.n_polys, pl.n_polys), dtype=torch.float64, device="cuda")
    fV, fE = pl.pack_grid(V[None])[0], pl.pack_grid(np.exp(-3 * V / 10)[None])[0]
    pl.varf(fV, "reference", out)
    pl.varf(fE, "gram", out)
elif which == "c5":
    x, w = gh(2001)
    keep = (w > 0) & (np.abs(x) <= 50)
    pl = Plan(1000, [x[keep]], [w[keep]])
    f = torch.randn(8192, pl.grid_size, dtype=torch.float64, device="cuda")
    c = pl.forward(f)
    pl.backward(c)
elif which == "c4":
    x, w = gh(256)
    pl = Plan(255, [x] * 3, [w] * 3, index_set="cube")
    f = torch.randn(1, pl.grid_size, dtype=torch.float64, device="cuda")
    c = pl.empty_coef(1)
    g = pl.empty_grid(1)
    for _ in range(2):
        pl.forward(f, c)
        pl.backward(c, g)
elif which == "gemm":
    from hermipy_b200.plan import dgemm
    M, N, K, cfg = (int(v) for v in sys.argv[2:6])
    os.environ["HB_GEMM_CFG"] = str(cfg)
    ev = lambda v: (v + 1) & ~1
    A = torch.randn((M, ev(K)), dtype=torch.float64, device="cuda")
    Bm = torch.randn((K, ev(N)), dtype=torch.float64, device="cuda")
    Cm = torch.empty((M, ev(N)), dtype=torch.float64, device="cuda")
    for _ in range(3):
        dgemm(M, N, K, A, ev(K), Bm, ev(N), Cm, ev(N))
    torch.cuda.synchronize()
    print("max err vs torch", float((Cm[:, :N] - A[:, :K] @ Bm[:, :N]).abs().max()))
elif which == "tables":
    # basis tables of the C5 rule (n = 1397 nodes, degree 1000) and the psi table of the C3 rule (717 nodes, degree 400)
    x, w = gh(2001)
    keep = (w > 0) & (np.abs(x) <= 50)
    for _ in range(2):
        pl = Plan(1000, [x[keep]], [w[keep]])
    from hermipy_b200.lib import gram_rule
    xg, wg, swg = gram_rule(801, 400)
    pl2 = Plan(400, [xg], [wg], sqrt_weights=[swg])
    pl2.varf(torch.ones(pl2.grid_size, dtype=torch.float64, device="cuda"), "gram")
elif which == "matvec":
    from hermipy_b200.resident import device_matvec
    n = 40401
    A = torch.randn((n, n), dtype=torch.float64, device="cuda")
    xv = torch.randn(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        yv = device_matvec(A, xv)
torch.cuda.synchronize()
print("done", which)
