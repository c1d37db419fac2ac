"""Warm, un-profiled split of the blocked Householder QR into its two phases, timed with CUDA events through the
block-level entry points (pla_qr_factor_block_f64 = cooperative block kernel + Gram/reflector build,
pla_qr_apply_block_f64 = W = Vx^T C, W2 = T^T W, C -= Vx W2).   usage: qr_phase_times.py M N [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parla_b200 import kernels as K

M, N = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nf = N - 1
ld = N + (N & 1)
g = torch.Generator(device="cuda").manual_seed(0)
W0 = torch.zeros(M, ld, dtype=torch.float64, device="cuda")
W0[:, :N] = torch.randn(M, N, dtype=torch.float64, device="cuda", generator=g)
ws = K.qr_block_workspace(W0.device, M, N)
best = None
for rep in range(reps):
    W = W0.clone()[:, :N]
    tau = torch.empty(nf, dtype=torch.float64, device="cuda")
    evs = []
    torch.cuda.synchronize()
    for blk, j0 in enumerate(range(0, nf, K.QR_BLOCK)):
        jb = min(K.QR_BLOCK, nf - j0)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        K.qr_factor_block(W, j0, j0, jb, tau[j0:j0 + jb], blk, N, ws)
        e[1].record()
        K.qr_apply_block(M, j0, jb, tau[j0:j0 + jb], W[:, j0 + jb:], N, ws)
        e[2].record()
        evs.append(e)
    torch.cuda.synchronize()
    f = sum(e[0].elapsed_time(e[1]) for e in evs)
    a = sum(e[1].elapsed_time(e[2]) for e in evs)
    if best is None or f + a < best[0] + best[1]:
        best = (f, a, [round(e[0].elapsed_time(e[1]), 3) for e in evs[:3]], [round(e[1].elapsed_time(e[2]), 3) for e in evs[:3]])
Wr = W0.clone()[:, :N]
K.geqrf(Wr, nf)
same = bool(torch.equal(torch.triu(Wr[:nf, :nf]), torch.triu(W[:nf, :nf])))
print(json.dumps({"M": M, "N": N, "factor_ms": best[0], "apply_ms": best[1], "first_blocks_factor_ms": best[2],
                  "first_blocks_apply_ms": best[3], "bit_identical_to_geqrf": same}))
