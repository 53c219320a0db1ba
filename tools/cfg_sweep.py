"""Developer script: time every tile configuration on the GEMM shapes of the bench workloads."""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermipy_b200 import _lib
lib = _lib.lib()
# private hook: hb_dev_dgemm uses cfg auto; sweep through env var HB_GEMM_CFG read by the library
def run(M,N,K,batch,am,shared_a=False,reps=10):
    ldA = (M+1)&~1 if am else (K+1)&~1
    rowsA = K if am else M
    A = torch.randn((1 if shared_a else batch)*rowsA*ldA, dtype=torch.float64, device="cuda")
    ldB=(N+1)&~1; B=torch.randn(batch*K*ldB, dtype=torch.float64, device="cuda")
    Cc=torch.empty(batch*M*ldB, dtype=torch.float64, device="cuda")
    st=C.c_void_p(torch.cuda.current_stream().cuda_stream)
    res={}
    for cfg in (-1,0,1,2,3,4,5,6,7,8,9,10,11):
        os.environ["HB_GEMM_CFG"]=str(cfg)
        def go():
            rc=lib.hb_dev_dgemm(M,N,K,batch,A.data_ptr(),ldA,0 if shared_a else rowsA*ldA,int(am),B.data_ptr(),ldB,K*ldB,Cc.data_ptr(),ldB,M*ldB,st)
            assert rc==0, rc
        go(); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): go()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/reps
        res[cfg]=ms
    fl=2.0*M*N*K*batch
    print("M=%6d N=%6d K=%4d b=%4d am=%d sa=%d : "%(M,N,K,batch,am,shared_a)+"  ".join("cfg%d %.1fus %.1fTF"%(c,ms*1e3,fl/ms*1e-9) for c,ms in res.items()))
B=17
run(B*401,201,401,1,0)            # c2 fwd stage 1
run(201,202,401,B,0,True)         # c2 fwd stage 2
run(B*201,401,201,1,0)            # c2 bwd stage 1
run(401,402,201,B,0,True)         # c2 bwd stage 2
run(65536,1001,1057,1,0,reps=3)   # c5 fwd
run(65536,256,256,1,0)            # c4 stage 1
run(256,256,256,256,0,True)       # c4 stage 2
run(256,65536,256,1,0)            # c4 stage 3
run(40602,40602,401,1,1,reps=2)   # c2 varf stage 2 shape (plain store)
