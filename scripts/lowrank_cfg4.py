"""BASELINE.json configs[3]: QB1 / QB2 + SVD1 randomized low-rank on a tall matrix with decaying
spectrum (A = (U sigma) V^T built on the device), rank k, 2 power iterations (RS1).
Default is the full 2^20 x 2^14, k = 512 (A = 137 GB); --m/--n/--k/--r shrink it."""
import argparse
import json
import math
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
    ap.add_argument("--m", type=int, default=1 << 20)
    ap.add_argument("--n", type=int, default=1 << 14)
    ap.add_argument("--k", type=int, default=512)
    ap.add_argument("--r", type=int, default=2048)
    ap.add_argument("--blk", type=int, default=128)
    ap.add_argument("--skip-qb2", action="store_true")
    a = ap.parse_args()
    m, n, k, r = a.m, a.n, a.k, a.r
    dev = "cuda"
    # warm-up on a toy problem: cuSOLVER/cuBLAS handles, module loading, workspace allocation
    Aw = torch.randn(4096, 512, dtype=torch.float64, device=dev)
    rla.SVD1(rla.QB2(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1)), 32, False))(Aw, 64, 0.0, 0, 1)
    rla.SVD1(rla.QB1(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1))))(Aw, 64, np.nan, 0, 1)
    del Aw
    g = torch.Generator(device=dev).manual_seed(0)
    torch.cuda.synchronize()
    t0 = time.time()
    U = rla.orth(torch.randn(m, r, dtype=torch.float64, device=dev, generator=g))
    V = rla.orth(torch.randn(n, r, dtype=torch.float64, device=dev, generator=g))
    sigma = torch.exp(-torch.arange(r, dtype=torch.float64, device=dev) / 100.0)
    torch.cuda.synchronize()
    t_orth = time.time() - t0
    orth_err = float(torch.linalg.norm(K.gemm(V, V, transa=True) - torch.eye(r, dtype=torch.float64, device=dev)))
    U.mul_(sigma)                                   # U <- U diag(sigma)
    A = K.gemm(U, V, transb=True)                   # (m x r)(r x n)
    del U
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    print(json.dumps({"build": "A=(U sigma)V^T", "m": m, "n": n, "r": r, "orth_s": round(t_orth, 3),
                      "orth_err_V": orth_err, "total_build_s": round(time.time() - t0, 3),
                      "mem_GB": round(torch.cuda.memory_allocated() / 1e9, 1)}), flush=True)
    tail = math.sqrt(float((sigma[k:] ** 2).sum()))          # best rank-k error
    normA = math.sqrt(float((sigma ** 2).sum()))

    def run(name, qb, tol):
        alg = rla.SVD1(qb)
        torch.cuda.synchronize(); t = time.time()
        Uh, s, Vh = alg(A, k, tol, 0, 1)
        torch.cuda.synchronize(); dt = time.time() - t
        s_true = sigma[:s.numel()]
        rec = {"alg": name, "m": m, "n": n, "k": int(s.numel()), "time_s": round(dt, 3),
               "TFLOPs_on_4_passes": round(4 * 2.0 * m * n * k / dt / 1e12, 2),
               "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1),
               "max_rel_sv_err": float(((s - s_true).abs() / s_true).max()),
               "orth_U": float(torch.linalg.norm(K.gemm(Uh, Uh, transa=True) - torch.eye(s.numel(), dtype=torch.float64, device=dev))),
               "passes_flops": 2.0 * m * n * k, "best_rank_k_rel_err": tail / normA}
        print(json.dumps(rec), flush=True)

    rs = rla.RS1(rla.SkOpGA(), 2, rla.orth, 1)
    run("SVD1(QB1(RF1(RS1(2 passes))))", rla.QB1(rla.RF1(rs)), np.nan)
    if not a.skip_qb2:
        run(f"SVD1(QB2(blk={a.blk}, overwrite_a=True))", rla.QB2(rla.RF1(rs), a.blk, True), 0.0)


if __name__ == "__main__":
    main()
