"""The two tensor-core products of one block-reflector application, timed alone at the first-block shapes of the
sketch QR:  W = Vx^T C (128 x nc, K = rows)  and  C -= Vx W2 (rows x nc, K = 128).   usage: time_qr_gemms.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parla_b200 import kernels as K

def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

for rows, n in [(8192, 2049), (16384, 4097), (16384, 2049)]:
    ld = n + (n & 1)
    nc = n - 128
    Cf = torch.randn(rows, ld, dtype=torch.float64, device="cuda")
    C = Cf[:, 128:n]
    Vx = torch.randn(rows, 128, dtype=torch.float64, device="cuda")
    ldw = nc + (nc & 1)
    W = torch.zeros(128, ldw, dtype=torch.float64, device="cuda")[:, :nc]
    W2 = torch.randn(128, ldw, dtype=torch.float64, device="cuda")[:, :nc]
    t1 = t(lambda: K.gemm(Vx, C, transa=True, out=W))
    t2 = t(lambda: K.gemm(Vx, W2, alpha=-1.0, beta=1.0, out=C))
    fl = 2.0 * rows * 128 * nc
    print(json.dumps({"rows": rows, "nc": nc, "w_eq_vxt_c_ms": t1, "w_tf": fl / t1 / 1e9, "c_minus_vx_w2_ms": t2,
                      "c_tf": fl / t2 / 1e9, "c_rw_gbs": 2 * rows * nc * 8 / t2 / 1e6}), flush=True)
