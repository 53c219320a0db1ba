"""Developer script: plain batched GEMM (hb_dev_dgemm, batch_lo = 1) per tile config, repeated, vs torch; for the
current library and for a round-1 build (tools/bin/libhermite_r1.so) if present."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200 import _lib
libs = {"current": _lib.lib()}
old = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libhermite_r1.so")
if os.path.exists(old):
    L = C.CDLL(old)
    L.hb_dev_dgemm.restype = C.c_int
    L.hb_dev_dgemm.argtypes = _lib.SIGNATURES["hb_dev_dgemm"][1]
    libs["round1"] = L
torch.manual_seed(1)
for (M, N, K, batch) in [(128, 4096, 128, 6), (128, 256, 128, 512), (401, 402, 201, 17), (128, 65536, 128, 2)]:
    A = torch.randn(batch, M, K + (K & 1), dtype=torch.float64, device="cuda")
    B = torch.randn(batch, K, N, dtype=torch.float64, device="cuda")
    want = torch.matmul(A[:, :, :K], B)
    for name, L in libs.items():
        res = {}
        for cfg in range(10):
            os.environ["HB_GEMM_CFG"] = str(cfg)
            bad = []
            for rep in range(6):
                Cm = torch.full((batch, M, N), float("nan"), dtype=torch.float64, device="cuda")
                rc = L.hb_dev_dgemm(M, N, K, batch, A.data_ptr(), A.stride(1), A.stride(0), 0, B.data_ptr(), N, K * N,
                                    Cm.data_ptr(), N, M * N, C.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0, rc
                torch.cuda.synchronize()
                bad.append(int(((Cm - want).abs() > 1e-9).sum() + (~torch.isfinite(Cm)).sum()))
            if any(bad):
                res[cfg] = bad
        os.environ["HB_GEMM_CFG"] = "-1"
        print("%s M %d N %d K %d batch %d: wrong entries per rep by cfg: %s" % (name, M, N, K, batch, res or "none"), flush=True)
