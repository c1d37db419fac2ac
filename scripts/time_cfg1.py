"""BASELINE.json configs[0] on the GPU: 2^16 x 500, SJLT k=8, d=4n, tol 1e-12 (reference: 2.9 s on 8 CPU cores)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parla_b200 as rla
rng = np.random.default_rng(0)
A = rng.standard_normal((65536, 500)); x0 = rng.standard_normal(500); b = A @ x0 + 0.1 * rng.standard_normal(65536)
Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
alg = rla.SPO(rla.SkOpSJ(8), 4, 'qr')
for i in range(3):
    alg(Ad, bd, 0.0, 1e-12, 100, i, logging=False)
torch.cuda.synchronize()
ts = []
for i in range(10):
    t = time.perf_counter(); x, _ = alg(Ad, bd, 0.0, 1e-12, 100, i, logging=False); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
x, log = alg(Ad, bd, 0.0, 1e-12, 100, 1, logging=True)
th = []
for i in range(3):
    t = time.perf_counter(); xh, _ = alg(A, b, 0.0, 1e-12, 100, i, logging=False); th.append(time.perf_counter() - t)
x_opt = np.linalg.lstsq(A, b, rcond=None)[0]
print(json.dumps({"cfg1": "SPO-qr SJLT 65536x500", "median_ms_device_resident": round(1e3 * sorted(ts)[5], 3),
                  "median_ms_host_buffers": round(1e3 * sorted(th)[1], 3), "iters": log.iters,
                  "phases_ms": {k: round(1e3 * getattr(log, "time_" + k), 3) for k in ("sketch", "factor", "presolve", "iterate")},
                  "rel_err_vs_lstsq": float(np.linalg.norm(x.cpu().numpy() - x_opt) / np.linalg.norm(x_opt))}))
