"""n x n factorisation behind mode='svd': cuSOLVER gesvd vs the Gram/eigh route (comps/preconditioning.py)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200.comps import preconditioning as rpc

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize(); return (time.time() - t0) / reps, r

for n in (2048, 4096):
    g = torch.Generator(device="cuda").manual_seed(n)
    R = torch.triu(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)) + 8 * torch.eye(n, dtype=torch.float64, device="cuda") * np.sqrt(n) / 8
    Rw = torch.linalg.qr(torch.randn(4 * n, n, dtype=torch.float64, device="cuda", generator=g))[1]      # R of a Gaussian sketch
    for name, X in (("triu+diag", R), ("R of 4n x n Gaussian", Rw)):
        t_eigh, _ = timed(lambda: torch.linalg.eigh(X.T @ X))
        rpc.FAST_SVD = False
        t_svd, ref = timed(lambda: rpc.svd_right_precond(X), reps=1)
        rpc.FAST_SVD = True
        t_fast, got = timed(lambda: rpc.svd_right_precond(X))
        Mr, Ur, sr, Vr = ref; Mg, Ug, sg, Vg = got
        print(json.dumps({"n": n, "matrix": name, "cond": float(sr[0] / sr[-1]), "gesvd_s": round(t_svd, 4), "eigh_only_s": round(t_eigh, 4),
                          "gram_route_s": round(t_fast, 4), "sigma_rel_err": float(((sg.sort(descending=True)[0] - sr).abs() / sr).max()),
                          "recon_err": float(torch.linalg.norm((Ug * sg) @ Vg - X) / torch.linalg.norm(X)),
                          "orth_U": float(torch.linalg.norm(Ug.T @ Ug - torch.eye(n, device="cuda", dtype=torch.float64))),
                          "precond_cond_minus_1": float(torch.linalg.cond(X @ Mg) - 1)}), flush=True)
