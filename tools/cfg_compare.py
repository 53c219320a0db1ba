"""Developer script: time one GEMM shape (A shared over `batch` B operands) under several tile configurations.
    python tools/cfg_compare.py M N K cfg [cfg ...]        (HB_BATCH=17 for the batched form)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import dgemm

M, N, K = (int(v) for v in sys.argv[1:4])
ev = lambda v: (v + 1) & ~1
nb = int(os.environ.get("HB_BATCH", "1"))
A = torch.randn((M, ev(K)), dtype=torch.float64, device="cuda")
B = torch.randn((nb, K, ev(N)), dtype=torch.float64, device="cuda")
C = torch.empty((nb, M, ev(N)), dtype=torch.float64, device="cuda")
_dgemm = dgemm
dgemm = lambda M, N, K, A, lda, B, ldb, C, ldc: _dgemm(M, N, K, A, lda, B, ldb, C, ldc, batch=nb, strideA=0,
                                                      strideB=B.stride(0), strideC=C.stride(0))
flush = torch.empty(64 << 20, dtype=torch.float64, device="cuda")
for cfg in sys.argv[4:]:
    os.environ["HB_GEMM_CFG"] = cfg
    for _ in range(3):
        dgemm(M, N, K, A, ev(K), B, ev(N), C, ev(N))
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dgemm(M, N, K, A, ev(K), B, ev(N), C, ev(N)); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("M %d N %d K %d cfg %s: median %.1f us  min %.1f us  (%.1f TFLOP/s at the median)" % (
        M, N, K, cfg, ts[len(ts) // 2], ts[0], 2.0 * nb * M * N * K / ts[len(ts) // 2] * 1e-6))
