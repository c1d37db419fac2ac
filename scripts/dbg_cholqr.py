import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import parla_b200 as rla
from parla_b200 import distla, kernels as K
g = torch.Generator(device="cuda").manual_seed(0)
Y = torch.randn(1 << 20, 512, dtype=torch.float64, device="cuda", generator=g)
for name, fn in (("cholqr2", lambda: distla._cholqr2(Y, None)), ("orth", lambda: rla.orth(Y)), ("householder", lambda: K.qr_economic(Y)[0])):
    Q = fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); Q = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if Q is None:
        print(name, "returned None"); continue
    e = float(torch.linalg.norm(K.gemm(Q, Q, transa=True) - torch.eye(512, dtype=torch.float64, device="cuda")))
    print(name, f"{1e3*dt:.1f} ms  |Q'Q-I| = {e:.2e}")
# pieces
G = K.gemm(Y, Y, transa=True); torch.cuda.synchronize()
for name, fn in (("gram gemm", lambda: K.gemm(Y, Y, transa=True)), ("cholesky_ex", lambda: torch.linalg.cholesky_ex(G, upper=True)),
                 ("trtri", lambda: K.trtri_upper(torch.linalg.cholesky_ex(G, upper=True)[0].contiguous())),
                 ("Y @ X", lambda: K.gemm(Y, G))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print("  ", name, f"{1e3*(time.perf_counter()-t0):.2f} ms")
