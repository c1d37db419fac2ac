"""FP64 DMMA GEMM at the shapes the drivers use (low-rank passes over A, Gram products, QR updates).
usage: time_gemm_shapes.py   (PARLA_B200_LIB=<other build> for an A/B on the same box)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parla_b200 import kernels as K

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

m, n, k = 1 << 18, 1 << 14, 512
A = torch.randn(m, n, dtype=torch.float64, device="cuda")
S = torch.randn(n, k, dtype=torch.float64, device="cuda")
Y = torch.empty(m, k, dtype=torch.float64, device="cuda")
Z = torch.empty(n, k, dtype=torch.float64, device="cuda")
G = torch.empty(k, k, dtype=torch.float64, device="cuda")
for name, fn, fl in [("Y=A S", lambda: K.gemm(A, S, out=Y), 2.0 * m * n * k),
                     ("Z=A^T Y", lambda: K.gemm(A, Y, transa=True, out=Z), 2.0 * m * n * k),
                     ("G=Y^T Y", lambda: K.gemm(Y, Y, transa=True, out=G), 2.0 * m * k * k)]:
    ms = t(fn)
    print(json.dumps({"gemm": name, "m": m, "n": n, "k": k, "ms": ms, "tf": fl / ms / 1e9, "lib": os.path.basename(os.environ.get("PARLA_B200_LIB", "libparla_b200.so"))}), flush=True)
