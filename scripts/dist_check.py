"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-sharded drivers must
reproduce the single-GPU results on the same global problem.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parla_b200 as rla                      # noqa: E402
from oracle import parla_oracle as orc        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    def report(name, err, tol):
        nonlocal ok
        good = err <= tol
        ok = ok and good
        if rank == 0:
            print(f"{'OK  ' if good else 'FAIL'} {name:58s} {err:.3e} (tol {tol:.0e})", flush=True)

    # ---------------- fused reduce + cross-GPU sum of the streaming pass (NVLink peer memory) vs pass + NCCL all-reduce
    from parla_b200 import kernels as K
    from parla_b200.parallel import peer_comm
    comm = peer_comm(dist.group.WORLD, dev)
    if rank == 0:
        print("peer exchange:", "ON (pla_stream_pass_peer_f64)" if comm is not None else "OFF (NCCL all-reduce)", flush=True)
    if comm is not None:
        for trial, (mm, nn) in enumerate([(5000, 500), (4096, 2048), (3000, 64), (2048, 4096), (3000, 64)]):
            g = torch.Generator(device=dev).manual_seed(50 * trial + rank)
            At = torch.randn(mm, nn, dtype=torch.float64, device=dev, generator=g)
            wt = torch.randn(nn, dtype=torch.float64, device=dev, generator=g)
            u1, u2 = torch.zeros(mm, dtype=torch.float64, device=dev), torch.zeros(mm, dtype=torch.float64, device=dev)
            z_ref = K.stream_pass(At, w=wt, u=u1, flags=K.PASS_DOT | K.PASS_AXPY)
            dist.all_reduce(z_ref)
            z_f = K.stream_pass(At, w=wt, u=u2, flags=K.PASS_DOT | K.PASS_AXPY, comm=comm)
            report(f"fused pass + peer sum {mm}x{nn} vs pass + NCCL all-reduce",
                   float(torch.linalg.vector_norm(z_f - z_ref) / torch.linalg.vector_norm(z_ref)), 1e-14)
            zs = [torch.empty_like(z_f) for _ in range(world)]
            dist.all_gather(zs, z_f)
            report(f"fused pass + peer sum {mm}x{nn} identical on every rank", max(float((z - zs[0]).abs().max()) for z in zs), 0.0)

    # ---------------- least squares: same global (A, b) on every rank, rows split evenly
    rng = np.random.default_rng(0)
    m, n = 4096 * world, 200
    A = rng.standard_normal((m, n)) * np.logspace(0, 2, n)
    b = A @ rng.standard_normal(n) + 0.1 * rng.standard_normal(m)
    Ad, bd = torch.from_numpy(A).to(dev), torch.from_numpy(b).to(dev)
    mine = slice(rank * (m // world), (rank + 1) * (m // world))
    Ash = rla.RowSharded.from_rank(Ad[mine].contiguous())
    bsh = rla.RowSharded.from_rank(bd[mine].contiguous())
    for gen_name, gen in (("SkOpGA", rla.SkOpGA()), ("SkOpSJ", rla.SkOpSJ(8))):
        for mode, delta in (("qr", 0.0), ("svd", 0.0), ("qr", 0.3)):
            x1, log1 = rla.SPO(gen, 4, mode)(Ad, bd, delta, 1e-12, 100, 5)
            xs, logs = rla.SPO(gen, 4, mode)(Ash, bsh, delta, 1e-12, 100, 5)
            err = float(torch.linalg.vector_norm(xs - x1) / torch.linalg.vector_norm(x1))
            report(f"SPO[{gen_name},{mode},delta={delta}] sharded vs 1 GPU, {logs.iters}/{log1.iters} its", err, 1e-10)
    S = orc.sjlt_operator(4 * n, m, np.random.default_rng(3), 8)          # replayed reference operator
    x_ref, _ = orc.SPO(lambda d, mm, r: S, 4, 'qr')(A, b, 0.0, 1e-12, 100, None)
    xs, _ = rla.SPO(lambda d, mm, r: S, 4, 'qr')(Ash, bsh, 0.0, 1e-12, 100, None)
    report("SPO[replayed scipy SJLT] sharded vs oracle", float(np.linalg.norm(xs.cpu().numpy() - x_ref) / np.linalg.norm(x_ref)), 1e-10)

    # ---------------- column-distributed QR of the replicated sketch (engages for n >= 2 * 128 * world)
    from parla_b200 import distla
    nq = 256 * world + 130
    g = torch.Generator(device=dev).manual_seed(11)
    W0 = torch.randn(4 * nq, nq + 2, dtype=torch.float64, device=dev, generator=g)
    dist.broadcast(W0, src=0)
    Wr, Wd = W0.clone(), W0.clone()
    K.geqrf(Wr[:, :nq + 1], nq)
    assert distla.geqrf_distributed_ok(Wd.shape[0], nq, dist.group.WORLD)
    distla.geqrf_distributed(Wd[:, :nq + 1], nq, dist.group.WORLD)
    Rr, Rd = torch.triu(Wr[:nq, :nq]), torch.triu(Wd[:nq, :nq])
    report("geqrf_distributed R vs single-GPU geqrf", float(torch.linalg.norm(Rd - Rr) / torch.linalg.norm(Rr)), 1e-13)
    report("geqrf_distributed Q^T b vs single-GPU geqrf", float(torch.linalg.norm(Wd[:nq, nq] - Wr[:nq, nq]) / torch.linalg.norm(Wr[:nq, nq])), 1e-13)
    Rall = [torch.empty_like(Rd) for _ in range(world)]
    dist.all_gather(Rall, Rd)
    report("geqrf_distributed R identical on every rank", max(float((Rall[r] - Rall[0]).abs().max()) for r in range(world)), 0.0)
    # the driver through it: a sketch wide enough for the distributed factorisation
    m2, n2 = 8192 * world, 256 * world + 64
    A3 = torch.randn(m2, n2, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(12))
    b3 = torch.randn(m2, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(13))
    mine3 = slice(rank * (m2 // world), (rank + 1) * (m2 // world))
    for mode in ("qr", "svd"):
        x1, l1 = rla.SPO(rla.SkOpSJ(8), 4, mode)(A3, b3, 0.0, 1e-12, 100, 5)
        xs, ls = rla.SPO(rla.SkOpSJ(8), 4, mode)(rla.RowSharded.from_rank(A3[mine3].contiguous()),
                                                rla.RowSharded.from_rank(b3[mine3].contiguous()), 0.0, 1e-12, 100, 5)
        report(f"SPO[SkOpSJ,{mode}] n={n2} (distributed QR) sharded vs 1 GPU, {ls.iters}/{l1.iters} its",
               float(torch.linalg.vector_norm(xs - x1) / torch.linalg.vector_norm(x1)), 1e-10)

    # ---------------- saddle-point systems (SPS2: LSQR; SPS1: PCG with SVD and Nystrom preconditioners)
    c = rng.standard_normal(n)
    cd = torch.from_numpy(c).to(dev)
    for name, mk, tol in (("SPS2[SkOpSJ]", lambda: rla.SPS2(rla.SkOpSJ(8), 4), 1e-12),
                          ("SPS2[SkOpGA]", lambda: rla.SPS2(rla.SkOpGA(), 4), 1e-12),
                          ("SPS1[SkOpSJ]", lambda: rla.SPS1(rla.SkOpSJ(8), 3), 1e-13)):
        x1, y1, log1 = mk()(Ad, bd, cd, 0.4, tol, 100, 9, logging=True)
        xs, ys, logs = mk()(Ash, bsh, cd, 0.4, tol, 100, 9, logging=True)
        report(f"{name} x sharded vs 1 GPU, {logs.iters}/{log1.iters} its",
               float(torch.linalg.vector_norm(xs - x1) / torch.linalg.vector_norm(x1)), 1e-10)
        report(f"{name} y (local rows) sharded vs 1 GPU",
               float(torch.linalg.vector_norm(ys - y1[mine]) / torch.linalg.vector_norm(y1)), 1e-10)
    for strat in ("left", "right"):
        def mk():
            a = rla.SPS1(orc.SkOpGA(), 0.85)
            a.nystrom_strategy = strat
            return a
        x1, y1, log1 = mk()(Ad, bd, cd, 0.4, 1e-13, 400, 9, logging=True)
        xs, ys, logs = mk()(Ash, bsh, cd, 0.4, 1e-13, 400, 9, logging=True)
        report(f"SPS1[Nystrom {strat}] x sharded vs 1 GPU, {logs.iters}/{log1.iters} its",
               float(torch.linalg.vector_norm(xs - x1) / torch.linalg.vector_norm(x1)), 1e-9)

    # ---------------- low rank: SVD1 over QB1 / QB2, numpy test matrices replayed (identical on both paths)
    A2 = orc.exponent_spectrum(2048 * world, 300, 120, np.random.default_rng(1), 8.0)
    A2d = torch.from_numpy(A2).to(dev)
    mine2 = slice(rank * 2048, (rank + 1) * 2048)
    for name, mk in (("QB1", lambda: rla.QB1(rla.RF1(rla.RS1(orc.SkOpGA(), 2, rla.orth, 1)))),
                     ("QB2", lambda: rla.QB2(rla.RF1(rla.RS1(orc.SkOpGA(), 2, rla.orth, 1)), 16, False)),
                     ("QB1 odd passes, native S", lambda: rla.QB1(rla.RF1(rla.RS1(rla.SkOpGA(), 3, rla.orth, 1))))):
        U1, s1, V1 = rla.SVD1(mk())(A2d, 40, 0.0 if name == "QB2" else np.nan, 0, 7)
        A2sh = rla.RowSharded.from_rank(A2d[mine2].contiguous())
        Us, ss, Vs = rla.SVD1(mk())(A2sh, 40, 0.0 if name == "QB2" else np.nan, 0, 7)
        report(f"SVD1[{name}] singular values sharded vs 1 GPU", float((ss - s1).abs().max() / s1[0]), 1e-10)
        ap1 = (U1 * s1) @ V1
        aps = (Us.local * ss) @ Vs
        report(f"SVD1[{name}] U s V^T (local rows) sharded vs 1 GPU", float(torch.linalg.norm(aps - ap1[mine2]) / torch.linalg.norm(ap1)), 1e-10)
        G = Us.local.T @ Us.local
        dist.all_reduce(G)
        report(f"SVD1[{name}] |U^T U - I| of the sharded U", float(torch.linalg.norm(G - torch.eye(G.shape[0], device=dev, dtype=G.dtype))), 1e-10)
    # symmetric EVD
    H = A2[:300 * 1].T @ A2[:300]
    nH = (H.shape[0] // world) * world
    H = 0.5 * (H + H.T)[:nH, :nH]
    Hd = torch.from_numpy(np.ascontiguousarray(H)).to(dev)
    V1, l1 = rla.EVD1(rla.QB1(rla.RF1(rla.RS1(orc.SkOpGA(), 2, rla.orth, 1))))(Hd, 20, np.nan, 5, 3)
    mineH = slice(rank * (nH // world), (rank + 1) * (nH // world))
    Vs, ls = rla.EVD1(rla.QB1(rla.RF1(rla.RS1(orc.SkOpGA(), 2, rla.orth, 1))))(rla.RowSharded.from_rank(Hd[mineH].contiguous()), 20, np.nan, 5, 3)
    report("EVD1 eigenvalues sharded vs 1 GPU", float((ls - l1).abs().max() / l1.abs().max()), 1e-10)

    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if float(flag) == 1.0 else "FAIL", f"world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 1.0 else 1)


if __name__ == "__main__":
    main()
