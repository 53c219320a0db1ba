import os, sys
sys.path.insert(0, "/root/repo")
import torch
from hermipy_b200.plan import dgemm
def run(M, N, K, batch, cfgs):
    A = torch.randn((batch, M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((batch, K, N), dtype=torch.float64, device="cuda")
    C = torch.empty((batch, M, N), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for cfg in cfgs:
        os.environ["HB_GEMM_CFG"] = str(cfg)
        for _ in range(3):
            dgemm(M, N, K, A, K, B, N, C, N, batch=batch, strideA=M * K, strideB=K * N, strideC=M * N)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dgemm(M, N, K, A, K, B, N, C, N, batch=batch, strideA=M * K, strideB=K * N, strideC=M * N)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print("M %d N %d K %d batch %d cfg %2d: %.1f us (min %.1f)" % (M, N, K, batch, cfg, ts[len(ts) // 2], ts[0]), flush=True)
run(128, 8192, 128, 2, (0, 2, 11, 10, 1))
run(8192, 128, 128, 2, (0, 2, 10, 11, 1))
run(128, 256, 128, 64, (0, 2, 10, 11, 1))
