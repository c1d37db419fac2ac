import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K
import parla_b200 as rla
def t(fn, name, reps=3):
    fn(); torch.cuda.synchronize()
    w=[]; g=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        t0=time.perf_counter(); e0.record(); fn(); e1.record(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
        w.append(t2-t0); g.append(e0.elapsed_time(e1)*1e-3); enq=t1-t0
    print(f"{name:40s} wall {min(w)*1e3:9.2f} ms   gpu {min(g)*1e3:9.2f} ms   enqueue {enq*1e3:9.2f} ms", flush=True)
m,n,k=262144,4096,256
gen=torch.Generator(device="cuda").manual_seed(0)
A=torch.randn(m,n,dtype=torch.float64,device="cuda",generator=gen)
Y=torch.randn(m,k,dtype=torch.float64,device="cuda",generator=gen)
S=torch.randn(n,k,dtype=torch.float64,device="cuda",generator=gen)
B=torch.randn(k,n,dtype=torch.float64,device="cuda",generator=gen)
t(lambda: K.qr_economic(Y), "qr_economic 262144x256")
W=Y.clone()
t(lambda: K.geqrf(W, k), "geqrf only 262144x256")
tau=K.geqrf(W,k)
t(lambda: K.orgqr(W,tau,k), "orgqr only")
t(lambda: K.qr_economic(S), "qr_economic 4096x256")
t(lambda: K.gemm(A,S), "gemm A@S")
t(lambda: K.gemm(A,Y,transa=True), "gemm A^T Y")
t(lambda: K.gemm(Y,A,transa=True), "gemm Y^T A")
t(lambda: torch.linalg.svd(B, full_matrices=False), "torch.linalg.svd 256x4096")
Y2=torch.randn(1<<20,512,dtype=torch.float64,device="cuda",generator=gen)
t(lambda: K.qr_economic(Y2), "qr_economic 2^20x512", reps=2)
