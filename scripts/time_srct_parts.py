"""Stage timing of the pruned two-level DCT (SRCTOperator._apply_factored) at the headline size, one column block."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parla_b200 as rla
from parla_b200 import kernels as K

m, n, d, w = 1 << 22, 2048, 8192, 512
A = torch.randn(m, n, dtype=torch.float64, device="cuda")
S = rla.srct_operator(d, m, 5)
m2 = S.choose_m2(m, d); m1 = m // m2
plan = S._plan(m2)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

Xp = torch.empty(m, w, dtype=torch.float64, device="cuda")
t_g, _ = timed(lambda: K.gather_rows_scale(A, S.perm, S.e, 0, w, Xp))
Y = torch.empty(2 * m2, m1 * w, dtype=torch.float64, device="cuda")
t_1, _ = timed(lambda: K.gemm(plan["F"], Xp.view(m2, m1 * w), out=Y))
kappa, g = plan["groups"][5]
ks, sg = plan["k_sorted"][:g], plan["sgn_sorted"][:g]
t_w, W2 = timed(lambda: K.srct_weights(ks, m, 0, m1, sgn=sg, with_sin=True))
t_2, Z = timed(lambda: K.gemm(W2, Y[2 * kappa - 1:2 * kappa + 1].view(2 * m1, w)))
ng = len(plan["groups"])
print(json.dumps({"m2": m2, "m1": m1, "groups": ng, "group_rows": g, "block_cols": w,
                  "gather_ms": round(t_g, 3), "level1_gemm_ms": round(t_1, 3),
                  "level1_TF": round(2.0 * 2 * m2 * m2 * m1 * w / t_1 / 1e9, 2),
                  "weights_ms_per_group": round(t_w, 3), "level2_gemm_ms_per_group": round(t_2, 3),
                  "level2_TF": round(2.0 * g * 2 * m1 * w / t_2 / 1e9, 2),
                  "est_total_ms_4_blocks": round(4 * (t_g + t_1 + ng * t_2) + ng * t_w, 1)}))
