"""BASELINE.json configs[2]: sketch-operator sweep, Gaussian (Philox-fused DMMA) vs SJLT (k=8), d = 4n,
m in {2^20, 2^22, 2^24}, n in {256, 1024, 4096}, restricted to m*n <= 2^34 (one GPU).  One JSON line
per (operator, m, n) with the roofline fraction of its bound (HBM for SJLT, FP64 tensor pipe for Gaussian)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K          # noqa: E402

HBM = 6455.6
DMMA = 37.1
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sorted(ts)[len(ts) // 2]


def main():
    only_small = "--small" in sys.argv
    for logm in (20, 22, 24):
        for n in (256, 1024, 4096):
            m, d = 1 << logm, 4 * n
            if m * n > (1 << 34) or (only_small and logm > 20):
                continue
            g = torch.Generator(device="cuda").manual_seed(0)
            A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
            b = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
            W = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")
            # SJLT: generation + plan + apply, and apply alone
            rows, signs = K.sjlt_generate(d, m, 8, 3)
            plan = K.SjltPlan(rows, signs, d)
            t_apply = timeit(lambda: plan.apply(A, 1 / math.sqrt(8), W, bvec=b, out_b=W[:, n]))
            t_plan = timeit(lambda: K.SjltPlan(K.sjlt_generate(d, m, 8, 3)[0], signs, d))
            bytes_alg = m * n * 8 + d * n * 8 + 5 * 8 * m
            print(json.dumps({"op": "sjlt_k8", "m": m, "n": n, "d": d, "apply_ms": round(t_apply * 1e3, 3),
                              "generate_plus_plan_ms": round(t_plan * 1e3, 3),
                              "GBps": round(bytes_alg / t_apply / 1e9, 1),
                              "frac_hbm_measured": round(bytes_alg / t_apply / 1e9 / HBM, 3)}), flush=True)
            del plan, rows, signs
            # Gaussian: cap the timed rows so one call stays ~<1 s; rate is independent of m
            mm = min(m, max(1 << 16, (1 << 36) // (d * n) // 4 * 4))
            t = timeit(lambda: K.sketch_gauss(A[:mm], d, 5, 1.0 / math.sqrt(d), W, bvec=b[:mm]), warm=1, reps=2)
            fl = 2.0 * d * mm * (n + 1)
            print(json.dumps({"op": "gauss_philox_dmma", "m": m, "n": n, "d": d, "rows_timed": mm,
                              "ms": round(t * 1e3, 3), "TFLOPs": round(fl / t / 1e12, 2),
                              "frac_dmma_measured": round(fl / t / 1e12 / DMMA, 3),
                              "full_m_estimate_s": round(t * m / mm, 3)}), flush=True)
            del A, b, W
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
