"""Developer script: time the resident-data paths at the BASELINE.json config sizes (run via gpurun)."""
import sys, os, time
import numpy as np, torch, scipy.special as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200.plan import Plan
from hermipy_b200 import _lib

def gh(n):
    x, w = sp.roots_hermitenorm(n); return x, w / np.sqrt(2 * np.pi)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

which = sys.argv[1:] or ["c5", "c2", "c4", "c3"]
# FP64 peak via cuBLAS
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
t = timeit(lambda: torch.matmul(a, b), reps=5); print("cuBLAS DGEMM 8192^3: %.2f ms %.2f TF" % (t, 2*8192**3/t*1e-9)); del a, b

if "c5" in which:
    x, w = gh(2001); keep = (w > 0) & (np.abs(x) <= 50); x, w = x[keep], w[keep]
    print("C5 nodes kept", len(x))
    pl = Plan(1000, [x], [w]); B = 65536
    f = torch.randn(B, pl.grid_size, dtype=torch.float64, device="cuda"); c = pl.empty_coef(B)
    t = timeit(lambda: pl.forward(f, c)); fl = 2.0*B*len(x)*1001
    print("C5 fwd: %.3f ms  %.2f TF  %.2f Gpts/s" % (t, fl/t*1e-9, B*len(x)/t*1e-6))
    g = pl.empty_grid(B)
    t = timeit(lambda: pl.backward(c, g))
    print("C5 bwd: %.3f ms  %.2f TF  %.2f Gpts/s" % (t, fl/t*1e-9, B*len(x)/t*1e-6))
    del f, c, g, pl

if "c2" in which:
    x, w = gh(401)
    for s in ("cube", "triangle"):
        pl = Plan(200, [x, x], [w, w], index_set=s)
        for B in (1, 17, 128):
            f = torch.randn(B, pl.grid_size, dtype=torch.float64, device="cuda"); c = pl.empty_coef(B); g = pl.empty_grid(B)
            tf = timeit(lambda: pl.forward(f, c), reps=20); tb = timeit(lambda: pl.backward(c, g), reps=20)
            print("C2 %s B=%d: fwd %.1f us (%.2f Gpts/s) bwd %.1f us (%.2f Gpts/s)" % (s, B, tf*1e3, B*401*401/tf*1e-6, tb*1e3, B*401*401/tb*1e-6))
        X, Y = np.meshgrid(x, x, indexing="ij")
        V = X**4/4 - X**2/2 + Y**4/4 - Y**2/2 + X*Y/2
        fV = pl.pack_grid(V[None])[0]; fE = pl.pack_grid(np.exp(-3*V/10)[None])[0]
        out = torch.empty((pl.n_polys, pl.n_polys), dtype=torch.float64, device="cuda")
        pl.set_timing(True)
        t = timeit(lambda: pl.varf(fV, "reference", out), reps=3, warm=1); k, p = pl.varf_info()
        print("C2 %s varf reference(poly): %.2f ms  kept=%d path=%d  %.1f GB/s written" % (s, t, k, p, pl.n_polys**2*8/t*1e-6))
        sym = (out - out.T).abs().max().item(); print("   symmetry err", sym)
        ref_poly = out.clone()
        t = timeit(lambda: pl.varf(fV, "gram", out), reps=3, warm=1); ms, fl = [(m_, f_) for t_, m_, f_ in pl.timings() if t_ == 22][0]
        print("C2 %s varf gram(poly): %.2f ms ; stage2 gemm %.2f ms %.2f TF(padded)" % (s, t, ms, fl/ms*1e-9))
        print("   gram vs reference rel err %.3e" % ((out-ref_poly).norm()/ref_poly.norm()).item())
        t = timeit(lambda: pl.varf(fE, "gram", out), reps=3, warm=1); ms, fl = [(m_, f_) for t_, m_, f_ in pl.timings() if t_ == 22][0]
        print("C2 %s varf gram(dense f): %.2f ms ; stage2 gemm %.2f ms %.2f TF(padded)" % (s, t, ms, fl/ms*1e-9))
        del out, ref_poly, pl

if "c4" in which:
    x, w = gh(256)
    t0 = time.time(); pl = Plan(255, [x]*3, [w]*3, index_set="cube"); print("C4 plan create %.1f s" % (time.time()-t0))
    f = torch.randn(1, pl.grid_size, dtype=torch.float64, device="cuda"); c = pl.empty_coef(1); g = pl.empty_grid(1)
    tf = timeit(lambda: pl.forward(f, c)); tb = timeit(lambda: pl.backward(c, g)); fl = 3*2*256.0**4
    print("C4 fwd %.3f ms (%.2f TF, %.2f Gpts/s) bwd %.3f ms (%.2f TF)" % (tf, fl/tf*1e-9, 256**3/tf*1e-6, tb, fl/tb*1e-9))
    pl.forward(f, c); pl.backward(c, g); print("   round trip rel err %.3e" % ((g-f).norm()/f.norm()).item())
    del f, c, g, pl

if "c3" in which:
    x, w = gh(801); keep = w > 0; x, w = x[keep], w[keep]; print("C3 nodes kept", len(x))
    pl = Plan(400, [x, x], [w, w], index_set="triangle")
    X, Y = np.meshgrid(x, x, indexing="ij"); V = X**4/4 - X**2/2 + Y**4/4 - Y**2/2 + X*Y/2
    fE = pl.pack_grid(np.exp(-3*V/10)[None])[0]
    out = torch.empty((pl.n_polys, pl.n_polys), dtype=torch.float64, device="cuda"); pl.set_timing(True)
    t = timeit(lambda: pl.varf(fE, "gram", out), reps=2, warm=1); ms, fl = [(m_, f_) for t_, m_, f_ in pl.timings() if t_ == 22][0]
    print("C3 triangle varf gram: %.1f ms ; stage2 gemm %.1f ms %.2f TF(padded full-square flops)" % (t, ms, fl/ms*1e-9))
