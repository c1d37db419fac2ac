import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import parla_b200 as rla
from parla_b200 import distla, kernels as K
m, n, k, r = (1 << 20, 16384, 512, 2048) if len(sys.argv) > 1 else (1 << 19, 8192, 512, 2048)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
U = rla.orth(torch.randn(m, r, dtype=torch.float64, device=dev, generator=g))
V = rla.orth(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
sigma = torch.exp(-torch.arange(r, dtype=torch.float64, device=dev) / 100.0)
U.mul_(sigma)
A = K.gemm(U, V, transb=True)
del U
orig = distla._cholqr2
def traced(Y, group):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Q = orig(Y, group)
    torch.cuda.synchronize()
    print(f"   _cholqr2 {tuple(Y.shape)} -> {'ok' if Q is not None else 'None'} in {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
    return Q
distla._cholqr2 = traced
alg = rla.SVD1(rla.QB1(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1))))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    alg(A, k, np.nan, 0, rep)
    torch.cuda.synchronize(); print(f"SVD1 total {time.perf_counter()-t0:.3f} s", flush=True)
os.environ["PLA_CHOLQR"] = "0"
torch.cuda.synchronize(); t0 = time.perf_counter()
alg(A, k, np.nan, 0, 5)
torch.cuda.synchronize(); print(f"SVD1 total (Householder orth) {time.perf_counter()-t0:.3f} s", flush=True)
