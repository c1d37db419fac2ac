"""Fused streaming pass on the narrow shapes (4 consumer groups per CTA): GB/s with the stage tags (default) or with the
ring rounded to a multiple of the group count (PLA_PASS_TAGS=0).  A is 4 GiB per n.  One JSON line per n."""
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
for n in (256, 500, 768, 1024):
    m = (1 << 29) // n
    A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    u = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    rec = {"n": n, "m": m, "tags": os.environ.get("PLA_PASS_TAGS", "1")}
    for name, kw in (("dot_axpy", dict(w=w, u=u, sa=1.0, su=-0.5, flags=3)), ("axpy_only", dict(u=u, flags=2))):
        for _ in range(2):
            K.stream_pass(A, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); K.stream_pass(A, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = sorted(ts)[3]
        rec[name + "_ms"] = round(1e3 * t, 3)
        rec[name + "_frac_hbm"] = round((m * n * 8 + 2 * m * 8) / t / 1e9 / peak, 3)
    print(json.dumps(rec), flush=True)
    del A, w, u
