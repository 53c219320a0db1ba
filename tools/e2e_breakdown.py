"""Developer script: where the end-to-end C2 step (host buffers in, host buffers out) spends its time.
    python tools/e2e_breakdown.py            (HB_PIPE_CHUNKS is read by the library at call time)"""
import os
import sys
import time

import numpy as np
import scipy.special as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200 import core  # noqa: E402

n, deg, B = 401, 200, 17
x, w = sp.roots_hermitenorm(n)
w = w / np.sqrt(2 * np.pi)
rng = np.random.default_rng(0)
Fh = torch.empty((B, n * n), dtype=torch.float64).pin_memory()
Fh.copy_(torch.from_numpy(rng.standard_normal((B, n * n))))
Fnp = Fh.numpy()
npol = core.iterator_size(2, deg, "cube")
Ch = torch.empty((B, npol), dtype=torch.float64).pin_memory().numpy()
Gh = torch.empty((B, n * n), dtype=torch.float64).pin_memory().numpy()
nodes, weights = [x, x], [w, w]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


fwd = lambda: core.transform_batch(deg, Fnp, nodes, weights, True, index_set="cube", out=Ch)
bwd = lambda: core.transform_batch(deg, Ch, nodes, weights, False, index_set="cube", out=Gh)
dF = torch.empty((B, n * n), dtype=torch.float64, device="cuda")
print("raw pinned H2D of the grid batch (%.1f MB): %.0f us" % (Fh.numel() * 8e-6, timeit(lambda: dF.copy_(Fh, non_blocking=True))))
Gt = torch.from_numpy(Gh)
print("raw pinned D2H of the grid batch: %.0f us" % timeit(lambda: Gt.copy_(dF, non_blocking=True)))
small = np.zeros((1, n * n))
so = np.zeros((1, npol))
print("one-function forward call (fixed costs + 1.3 MB): %.0f us" % timeit(
    lambda: core.transform_batch(deg, small, nodes, weights, True, index_set="cube", out=so)))
for chunks in (os.environ.get("HB_PIPE_CHUNKS", "default"),):
    print("chunks=%s forward %.0f us, inverse %.0f us" % (chunks, timeit(fwd), timeit(bwd)))
