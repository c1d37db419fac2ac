"""Time and check pla_geqrf_f64 / pla_orgqr_f64 at the sketch-QR shapes of the BASELINE configs.
usage: python scripts/time_qr.py [reps]      (env PLA_QR_COOP=0 -> the round-1 panel path; PLA_QR_RPC=rows per CTA)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from parla_b200 import kernels as K

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N) in [(2000, 501), (4096, 1025), (8192, 2049), (10000, 2001), (16384, 4097), (300, 64), (1000, 130)]:
    # the drivers keep the sketch in a buffer with an even row pitch (least_squares._sketch); PLA_QR_PITCH=odd times
    # a contiguous d x (n + 1) buffer instead (8-byte operand copies in the trailing updates)
    ld = N if os.environ.get("PLA_QR_PITCH") == "odd" else N + (N & 1)
    W0f = torch.zeros(M, ld, dtype=torch.float64, device="cuda")
    W0f[:, :N] = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
    W0 = W0f[:, :N]
    nf = N - 1
    # correctness vs torch (cuSOLVER) R, rows sign-normalised
    W = W0f.clone()[:, :N]
    tau = K.geqrf(W, nf)
    R = torch.triu(W[:nf, :nf])
    Rref = torch.linalg.qr(W0[:, :nf], mode='r')[1]
    sg = torch.sign(R.diagonal()) * torch.sign(Rref.diagonal())
    errR = float(torch.linalg.norm(R - sg[:, None] * Rref) / torch.linalg.norm(Rref))
    # Q^T b in the last column: |Q^T b| = |proj|
    x = K.trsv_upper(W[:nf, :nf], W[:nf, nf].contiguous())
    xref = torch.linalg.lstsq(W0[:, :nf], W0[:, nf:nf + 1]).solution[:, 0]
    errx = float(torch.linalg.norm(x - xref) / torch.linalg.norm(xref))
    ts = []
    for i in range(reps):
        W = W0f.clone()[:, :N]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.geqrf(W, nf); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    flops = 2.0 * M * nf * nf - 2.0 * nf ** 3 / 3
    rec = {"op": "geqrf", "M": M, "N": N, "ms_best": min(ts), "ms_median": sorted(ts)[len(ts) // 2],
           "tflops": flops / min(ts) / 1e9, "errR_vs_cusolver": errR, "err_lstsq": errx,
           "coop": os.environ.get("PLA_QR_COOP", "1"), "rpc": os.environ.get("PLA_QR_RPC", "64"), "pitch": ld}
    print(json.dumps(rec), flush=True)
# orth (geqrf + orgqr) of a tall-skinny block, as in the low-rank path
for (M, N) in [(1 << 17, 128), (1 << 20, 512)]:
    Y = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
    Q, R = K.qr_economic(Y)
    eo = float(torch.linalg.norm(Q.T @ Q - torch.eye(N, dtype=torch.float64, device="cuda")))
    er = float(torch.linalg.norm(Q @ R - Y) / torch.linalg.norm(Y))
    ts = []
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); K.qr_economic(Y); torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(json.dumps({"op": "qr_economic", "M": M, "N": N, "ms_best": min(ts), "orth_err": eo, "recon_err": er}), flush=True)
