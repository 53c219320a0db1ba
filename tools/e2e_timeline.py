"""Developer script: per-chunk timeline of the host-pointer batch transform (run with HB_PIPE_DEBUG=1)."""
import os
import sys

import numpy as np
import scipy.special as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200 import core  # noqa: E402

n, deg, B = 401, 200, 17
x, w = sp.roots_hermitenorm(n)
w = w / np.sqrt(2 * np.pi)
Fh = torch.empty((B, n * n), dtype=torch.float64).pin_memory()
Fh.copy_(torch.from_numpy(np.random.default_rng(0).standard_normal((B, n * n))))
npol = core.iterator_size(2, deg, "cube")
Ch = torch.empty((B, npol), dtype=torch.float64).pin_memory().numpy()
Gh = torch.empty((B, n * n), dtype=torch.float64).pin_memory().numpy()
for rep in range(4):
    core.transform_batch(deg, Fh.numpy(), [x, x], [w, w], True, index_set="cube", out=Ch)
    core.transform_batch(deg, Ch, [x, x], [w, w], False, index_set="cube", out=Gh)
one_f = torch.empty((1, n * n), dtype=torch.float64).pin_memory().numpy()
one_c = torch.empty((1, npol), dtype=torch.float64).pin_memory().numpy()
for rep in range(3):
    core.transform_batch(deg, one_f, [x, x], [w, w], True, index_set="cube", out=one_c)
