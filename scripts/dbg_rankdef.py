"""debug: rank-deficient SPO(svd) on the device vs the fixture"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import parla_b200 as rla
from parla_b200 import kernels as K
from parla_b200.comps import preconditioning as rpc
from tests.helpers import load_golden, Replay
np.set_printoptions(linewidth=200, precision=4)
for name in ("spo_gauss_svd_rankdef_seed1", "spo_gauss_svd_rankdef_seed4"):
    fx = load_golden(name)
    A, b, S = fx["A"], fx["b"], fx["S"]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x, log = rla.SPO(Replay(S), 3, 'svd')(dev(A), dev(b), 0.0, 1e-12, 100, None)
    x = x.cpu().numpy()
    print(name, "dx", np.linalg.norm(x - fx["x"]) / np.linalg.norm(fx["x"]), "errors", log.errors, "ref", fx["errors"])
    W = dev(np.hstack([S @ A, (S @ b)[:, None], np.zeros((30, 1))]))[:, :11]
    K.geqrf(W, 10)
    R = torch.triu(W[:10, :10])
    print(" svd(R_qr):", torch.linalg.svdvals(R).cpu().numpy())
    print(" svd(A_ske):", np.linalg.svd(S @ A, compute_uv=False))
    M, U, s, Vh = rpc.svd_right_precond(R)
    print(" rank", s.numel())
