"""Fused streaming pass (u <- sa A w + su u, z = A^T u, |u|^2 in one read of A) over the column counts of the
verdict's list; A is ~16 GiB for every n (>> L2).  One JSON line per n."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K

peak = 6455.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device="cuda").manual_seed(0)
for n in (256, 500, 1024, 2048, 2049, 4096, 8192, 16384):
    m = (1 << 31) // n
    A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    u = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    for _ in range(2):
        K.stream_pass(A, w=w, u=u, sa=1.0, su=-0.5, flags=3)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.stream_pass(A, w=w, u=u, sa=1.0, su=-0.5, flags=3); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = sorted(ts)[2]
    reads = 2 if (n > K.PASS_MAX_N or (n % 2 and n > K.PASS_MAX_N // 2)) else 1     # wide matrices: column blocks, 2 sweeps
    alg = m * n * 8 + 2 * m * 8
    print(json.dumps({"kernel": "stream_pass_fused", "m": m, "n": n, "ms": round(1e3 * t, 3), "reads_of_A": reads,
                      "GBps_algorithmic(one read)": round(alg / t / 1e9, 1), "frac_hbm_measured": round(alg / t / 1e9 / peak, 3),
                      "frac_of_the_reads_actually_made": round(reads * alg / t / 1e9 / peak, 3)}), flush=True)
    del A, w, u
# explicit triangular inverse (R^-1 preconditioner) and its pieces
for n in (2048, 4096):
    R = torch.triu(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)) + 60 * torch.eye(n, dtype=torch.float64, device="cuda")
    K.trtri_upper(R); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); X = K.trtri_upper(R); e1.record(); torch.cuda.synchronize()
    err = float(torch.linalg.norm(X @ R - torch.eye(n, dtype=torch.float64, device="cuda")))
    print(json.dumps({"kernel": "trtri_upper", "n": n, "ms": round(e0.elapsed_time(e1), 3), "|XR-I|": err}), flush=True)
