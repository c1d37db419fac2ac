"""Tiny invocation of every kernel family, meant to run under compute-sanitizer
(memcheck / racecheck) on a B200:  compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parla_b200 as rla                      # noqa: E402
from parla_b200 import kernels as K           # noqa: E402

rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for (m, n) in ((300, 40), (513, 77), (700, 600), (64, 2050)):
    A, w, u = rng.standard_normal((m, n)), rng.standard_normal(n), rng.standard_normal(m)
    ud = dev(u)
    z = K.stream_pass(dev(A), w=dev(w), u=ud, sa=0.5, su=-1.0, flags=3).cpu().numpy()
    ur = 0.5 * (A @ w) - u
    assert np.allclose(ud.cpu().numpy(), ur) and np.allclose(z[:n], A.T @ ur)
R = np.linalg.qr(rng.standard_normal((150, 70)))[1]
for tr in (False, True):
    x = K.trsv_upper(dev(R), dev(np.ones(70)), trans=tr).cpu().numpy()
    assert np.allclose((R.T if tr else R) @ x, 1.0)
assert np.allclose(K.trtri_upper(dev(R)).cpu().numpy() @ R, np.eye(70), atol=1e-10)
for ta, tb in ((0, 0), (1, 0), (0, 1), (1, 1)):
    Am, Bm = rng.standard_normal((90, 150) if ta else (150, 90)), rng.standard_normal((70, 90) if tb else (90, 70))
    C = K.gemm(dev(Am), dev(Bm), transa=bool(ta), transb=bool(tb)).cpu().numpy()
    assert np.allclose(C, (Am.T if ta else Am) @ (Bm.T if tb else Bm))
Y = rng.standard_normal((600, 40))
Q, Rr = K.qr_economic(dev(Y))
assert np.allclose((Q @ Rr).cpu().numpy(), Y)
A = rng.standard_normal((900, 30)); b = rng.standard_normal(900)
for gen in (rla.SkOpSJ(8), rla.SkOpGA()):
    for mode, delta in (("qr", 0.0), ("svd", 0.2)):
        x, log = rla.SPO(gen, 4, mode)(dev(A), dev(b), delta, 1e-12, 60, 1)
        ref = np.linalg.lstsq(np.vstack([A, np.sqrt(delta) * np.eye(30)]), np.concatenate([b, np.zeros(30)]), rcond=None)[0]
        assert np.linalg.norm(x.cpu().numpy() - ref) < 1e-9 * np.linalg.norm(ref)
y, _ = rla.SPU1(rla.SkOpSJ(8), 4)(dev(A), dev(rng.standard_normal(30)), 1e-12, 60, 1)
U, s, Vh = rla.SVD1(rla.QB2(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1)), 8, False))(dev(A), 16, 0.0, 0, 1)
c = rng.standard_normal(30)
G = A.T @ A + 0.2 * np.eye(30)
for gen in (rla.SkOpSJ(8), rla.SkOpGA()):
    for alg in (rla.SPS2(gen, 4), rla.SPS1(gen, 4)):
        x, yv, log = alg(dev(A), dev(b), dev(c), 0.2, 1e-12, 60, 1, logging=True)
        ref = np.linalg.solve(G, A.T @ b - c)
        assert np.linalg.norm(x.cpu().numpy() - ref) < 1e-8 * np.linalg.norm(ref)
nys = rla.SPS1(rla.SkOpGA(), 0.8)
x, yv, log = nys(dev(A), dev(b), dev(c), 0.2, 1e-12, 60, 1, logging=True)
assert np.linalg.norm(x.cpu().numpy() - ref) < 1e-8 * np.linalg.norm(ref)
# ---- round 2 additions: cooperative block QR (all-to-all and fan-out exchanges, partial last block), batched trtri
#      levels, wide / 512-thread streaming pass, GEMM with beta folded into the accumulators, oversized SJLT bucket,
#      rank-truncated SVD mode, CholeskyQR2
for (m, n) in ((2600, 300), (1300, 141)):
    Y = rng.standard_normal((m, n))
    Q, Rr = K.qr_economic(dev(Y))
    assert np.allclose((Q @ Rr).cpu().numpy(), Y)
R2 = np.linalg.qr(rng.standard_normal((400, 200)))[1]
assert np.allclose(K.trtri_upper(dev(R2)).cpu().numpy() @ R2, np.eye(200), atol=1e-9)
for (m, n) in ((40, 8448), (60, 8192), (50, 4099)):
    A2, w2, u2 = rng.standard_normal((m, n)), rng.standard_normal(n), rng.standard_normal(m)
    ud = dev(u2)
    z = K.stream_pass(dev(A2), w=dev(w2), u=ud, sa=0.5, su=-1.0, flags=3).cpu().numpy()
    ur = 0.5 * (A2 @ w2) - u2
    assert np.allclose(ud.cpu().numpy(), ur) and np.allclose(z[:n], A2.T @ ur)
Cm = rng.standard_normal((150, 70)); Am = rng.standard_normal((150, 90)); Bm = rng.standard_normal((90, 70))
Cd = dev(Cm)
K.gemm(dev(Am), dev(Bm), alpha=-1.0, beta=1.0, out=Cd)
assert np.allclose(Cd.cpu().numpy(), Cm - Am @ Bm)
rows = np.full((9000, 1), 2, dtype=np.int32); signs = np.ones((9000, 1), dtype=np.int8)
plan = K.SjltPlan(dev(rows), dev(signs), 8, validate=True)
outp = torch.empty(8, 5, dtype=torch.float64, device="cuda")
A3 = rng.standard_normal((9000, 5))
plan.apply(dev(A3), 1.0, outp)
assert np.allclose(outp.cpu().numpy()[2], A3.sum(axis=0))
Ald = rng.standard_normal((400, 6)) @ rng.standard_normal((6, 12))
x, log = rla.SAP2(rla.SkOpGA(), 3)(dev(Ald), dev(Ald @ rng.standard_normal(12)), 0.0, 1e-12, 50, 2)
assert np.all(np.isfinite(x.cpu().numpy()))
Yt = torch.randn(1 << 15, 24, dtype=torch.float64, device="cuda")
Qt = rla.orth(Yt)
assert float(torch.linalg.norm(Qt.T @ Qt - torch.eye(24, dtype=torch.float64, device="cuda"))) < 1e-12
# fused reduce + peer exchange (world = 1: the exchange buffer is this process's own), update-form GEMM on an even pitch
from parla_b200.parallel import PeerComm
comm = PeerComm(None, torch.device("cuda", 0), 8193)
for (m, n) in ((300, 64), (257, 2049), (300, 64)):
    A4, w4 = dev(rng.standard_normal((m, n))), dev(rng.standard_normal(n))
    ua, ub = torch.zeros(m, dtype=torch.float64, device="cuda"), torch.zeros(m, dtype=torch.float64, device="cuda")
    assert torch.equal(K.stream_pass(A4, w=w4, u=ua, flags=3), K.stream_pass(A4, w=w4, u=ub, flags=3, comm=comm))
torch.cuda.synchronize()
comm.close()
Cbuf = torch.zeros(300, 262, dtype=torch.float64, device="cuda")
Cm2 = rng.standard_normal((300, 261)); Cbuf[:, :261] = dev(Cm2)
Am2, Bm2 = rng.standard_normal((300, 128)), rng.standard_normal((128, 261))
Bbuf = torch.zeros(128, 262, dtype=torch.float64, device="cuda"); Bbuf[:, :261] = dev(Bm2)
K.gemm(dev(Am2), Bbuf[:, :261], alpha=-1.0, beta=1.0, out=Cbuf[:, :261])
assert np.allclose(Cbuf[:, :261].cpu().numpy(), Cm2 - Am2 @ Bm2) and float(Cbuf[:, 261].abs().sum()) == 0.0
torch.cuda.synchronize()
print("sanitize_smoke OK")
