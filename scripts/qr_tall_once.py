import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K
g = torch.Generator(device="cuda").manual_seed(0)
Y = torch.randn(1 << 20, 512, dtype=torch.float64, device="cuda", generator=g)
Q, R = K.qr_economic(Y); torch.cuda.synchronize()
Q, R = K.qr_economic(Y); torch.cuda.synchronize()
print("orth err", float(torch.linalg.norm(K.gemm(Q, Q, transa=True) - torch.eye(512, dtype=torch.float64, device="cuda"))))
