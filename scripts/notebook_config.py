"""The one timing the reference publishes (notebooks/least_squares/sap1_vs_lapack.ipynb:61-124):
SPO(sketch, sampling_factor=5) on a 100 000 x 2 000 fp64 matrix with a linear / log-spaced spectrum (cond 1e5),
b with 95 % of its mass in range(A), tol 1e-12, iter_lim = n.  Published (unstated CPU): SRCT 12.0-13.8 s,
SJLT 8.7-9.5 s, LAPACK lstsq 19.7-21.5 s.  Here: the same problem built on the device, the same call, timed
with the operator generation inside (as in the notebook), device-resident A and from host buffers.

    python scripts/notebook_config.py [--rows 100000 --cols 2000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parla_b200 as rla                      # noqa: E402
from parla_b200 import kernels as K           # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--cols", type=int, default=2000)
    ap.add_argument("--host", action="store_true", help="also time the call with numpy inputs (uploads inside)")
    a = ap.parse_args()
    m, n, kappa, prop = a.rows, a.cols, 1e5, 0.95
    g = torch.Generator(device="cuda").manual_seed(0)
    U = K.qr_economic(torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g))[0]
    Vt = K.qr_economic(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g))[0].T.contiguous()
    for spec_name, spec in (("linear", np.linspace(kappa ** 0.5, kappa ** -0.5, num=n)),
                            ("log", np.logspace(np.log10(kappa) / 2, -np.log10(kappa) / 2, num=n))):
        s = torch.from_numpy(spec).cuda()
        A = K.gemm((U * s).contiguous(), Vt)
        b0 = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        br = U @ (U.T @ b0)
        bo = b0 - br
        b = prop * br * (spec.mean() / br.norm()) + (1 - prop) * bo * (spec.mean() / bo.norm())
        x_opt = (Vt.T / s) @ (U.T @ b)
        for name, gen in (("srct", rla.srct_operator), ("sjlt", rla.sjlt_operator), ("gauss", rla.gaussian_operator)):
            sap = rla.SPO(gen, sampling_factor=5)
            times = []
            for seed in (11, 12, 13):
                torch.cuda.synchronize()
                t0 = time.time()
                x, log = sap(A, b, 0.0, 1e-12, n, np.random.default_rng(seed))
                torch.cuda.synchronize()
                times.append(time.time() - t0)
            rec = {"config": f"notebook sap1_vs_lapack: {m}x{n}, {spec_name} spectrum cond 1e5, sf=5, tol 1e-12",
                   "sketch": name, "solve_s": [round(t, 4) for t in times], "iters": log.iters,
                   "phases_s": {"sketch": round(log.time_sketch, 4), "factor": round(log.time_factor, 4),
                                "presolve": round(log.time_presolve, 4), "iterate": round(log.time_iterate, 4)},
                   "rel_err_vs_x_opt": float((x - x_opt).norm() / x_opt.norm())}
            if a.host:
                Ah, bh = A.cpu().numpy(), b.cpu().numpy()
                t0 = time.time()
                xh, _ = sap(Ah, bh, 0.0, 1e-12, n, np.random.default_rng(14))
                rec["solve_from_host_s"] = round(time.time() - t0, 4)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
