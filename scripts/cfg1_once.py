"""BASELINE.json configs[0] once (after warm-up), for ncu launch lists."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parla_b200 as rla
rng = np.random.default_rng(0)
A = rng.standard_normal((65536, 500)); x0 = rng.standard_normal(500); b = A @ x0 + 0.1 * rng.standard_normal(65536)
Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
alg = rla.SPO(rla.SkOpSJ(8), 4, 'qr')
for i in range(2):
    alg(Ad, bd, 0.0, 1e-12, 100, i, logging=False)
torch.cuda.synchronize()
