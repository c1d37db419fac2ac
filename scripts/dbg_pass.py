import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K
g = torch.Generator(device="cuda").manual_seed(0)
for n in (1024, 2048, 2049, 4096, 8192):
    for m in (20000, (1 << 31) // n):
        A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
        w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        u = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        u0 = u.clone()
        zss = K.stream_pass(A, w=w, u=u, sa=1.0, su=-0.5, flags=3)
        torch.cuda.synchronize()
        ref = A[:1000] @ w - 0.5 * u0[:1000]
        print(n, m, "ok", float((u[:1000] - ref).abs().max()), flush=True)
        del A, w, u
