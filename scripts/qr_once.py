"""One warm-up + one geqrf of an M x N sketch-shaped matrix (for ncu launch lists).  usage: qr_once.py M N"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K
M, N = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(0)
ld = N + (N & 1)                      # even row pitch, as least_squares._sketch allocates the sketch
W0 = torch.zeros(M, ld, dtype=torch.float64, device="cuda")
W0[:, :N] = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
for _ in range(2):
    W = W0.clone()[:, :N]; K.geqrf(W, N - 1); torch.cuda.synchronize()
