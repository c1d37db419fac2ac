"""Kernel-level timings on one B200 (CUDA events, warm-up, inputs >> L2).  Prints one JSON line per
kernel; used to fill profiles/ and DESIGN.md.  Usage: python scripts/microbench.py [--m 4194304 --n 2048]"""
import argparse
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parla_b200 import kernels as K          # noqa: E402
import parla_b200 as rla                      # noqa: E402


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=1 << 22)
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--skip-gauss", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    m, n, d = a.m, a.n, 4 * a.n
    peak = 6455.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    only = set(a.only.split(",")) if a.only else None
    want = lambda k: only is None or k in only
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    u = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    bytesA = m * n * 8
    out = {}

    def report(name, t, tmin, alg_bytes=None, flops=None, extra=None):
        rec = {"kernel": name, "m": m, "n": n, "ms": round(t * 1e3, 4), "ms_min": round(tmin * 1e3, 4)}
        if alg_bytes:
            rec["GBps"] = round(alg_bytes / t / 1e9, 1)
            rec["frac_hbm_measured"] = round(alg_bytes / t / 1e9 / peak, 3)
        if flops:
            rec["TFLOPs"] = round(flops / t / 1e12, 2)
        if extra:
            rec.update(extra)
        print(json.dumps(rec), flush=True)

    if want("dmma"):
        from parla_b200 import _lib
        lib = _lib.load()
        sink = torch.zeros(1, dtype=torch.float64, device="cuda")
        for cps in (1, 2, 4):
            iters = 20000
            fn = lambda: lib.pla_dmma_probe(iters, cps, sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
            t, tm = timeit(fn, warm=1, reps=3)
            fl = 2 * 256 * 8 * iters * 8 * lib.pla_num_sms() * cps
            report(f"dmma_probe ctas/sm={cps}", t, tm, flops=fl)
    if want("pass"):
        t, tm = timeit(lambda: K.stream_pass(A, w=w, u=u, sa=1.0, su=-0.5, flags=3))
        report("stream_pass_fused(dot+axpy)", t, tm, bytesA + 2 * m * 8)
        t, tm = timeit(lambda: K.stream_pass(A, w=w, u=u, sa=1.0, su=-0.5, flags=1))
        report("stream_pass(dot only)", t, tm, bytesA + 2 * m * 8)
        t, tm = timeit(lambda: K.stream_pass(A, u=u, flags=2))
        report("stream_pass(axpy only)", t, tm, bytesA + m * 8)
        t, tm = timeit(lambda: A.sum())
        report("torch.sum(A) [read-only reference]", t, tm, bytesA)
    if want("trsv"):
        R = torch.triu(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)) + 50 * torch.eye(n, dtype=torch.float64, device="cuda")
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        t, tm = timeit(lambda: K.trsv_upper(R, w, trans=False, out=x), reps=20)
        report("trsv_upper N", t, tm)
        t, tm = timeit(lambda: K.trsv_upper(R, w, trans=True, out=x), reps=20)
        report("trsv_upper T", t, tm)
    if want("sjlt"):
        rows, signs = K.sjlt_generate(d, m, 8, 3)
        torch.cuda.synchronize()
        t, tm = timeit(lambda: K.sjlt_generate(d, m, 8, 3))
        report("sjlt_generate", t, tm)
        t, tm = timeit(lambda: K.SjltPlan(rows, signs, d), reps=3)
        report("sjlt_plan", t, tm)
        plan = K.SjltPlan(rows, signs, d)
        W = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")
        t, tm = timeit(lambda: plan.apply(A, 1 / math.sqrt(8), W, bvec=u, out_b=W[:, n]), reps=3)
        report("sjlt_apply", t, tm, bytesA + d * n * 8 + 5 * 8 * m)
        del plan, rows, signs
    if want("qr"):
        W = torch.randn(d, n + 1, dtype=torch.float64, device="cuda", generator=g)
        W0 = W.clone()
        def qr():
            W.copy_(W0)
            K.geqrf(W, n)
        t, tm = timeit(qr, warm=1, reps=3)
        report("geqrf d x (n+1)", t, tm, flops=2 * d * n * n - 2 * n ** 3 / 3, extra={"d": d})
    if want("gemm"):
        k = 512
        Sm = torch.randn(n, k, dtype=torch.float64, device="cuda", generator=g)
        mm = min(m, 1 << 20)
        Y = torch.empty(mm, k, dtype=torch.float64, device="cuda")
        t, tm = timeit(lambda: K.gemm(A[:mm], Sm, out=Y), reps=3)
        report("gemm NN (A S)", t, tm, flops=2 * mm * n * k, extra={"k": k, "mm": mm})
        Z = torch.empty(n, k, dtype=torch.float64, device="cuda")
        t, tm = timeit(lambda: K.gemm(A[:mm], Y, transa=True, out=Z), reps=3)
        report("gemm TN (A^T Q)", t, tm, flops=2 * mm * n * k, extra={"k": k, "mm": mm})
    if want("gauss") and not a.skip_gauss:
        mm = min(m, 1 << 18)
        W = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")
        t, tm = timeit(lambda: K.sketch_gauss(A[:mm], d, 5, 1.0, W, bvec=u[:mm]), warm=1, reps=3)
        report("sketch_gauss", t, tm, flops=2 * d * mm * (n + 1), extra={"mm": mm, "d": d})
    if want("spo"):
        x0 = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        b, _ = K.matvec(A, x0)
        b += 0.1 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        for name, gen in (("sjlt", rla.SkOpSJ(8)),):
            x, log = rla.SPO(gen, 4, 'qr')(A, b, 0.0, 1e-12, 100, 3)
            x, log = rla.SPO(gen, 4, 'qr')(A, b, 0.0, 1e-12, 100, 3)
            tot = log.time_sketch + log.time_factor + log.time_presolve + log.time_iterate
            print(json.dumps({"solve": "SPO-qr-" + name, "m": m, "n": n, "total_s": round(tot, 4),
                              "sketch": round(log.time_sketch, 4), "factor": round(log.time_factor, 4),
                              "presolve": round(log.time_presolve, 4), "iterate": round(log.time_iterate, 4),
                              "iters": log.iters, "passes": log.passes_over_A,
                              "ms_per_iter": round(1e3 * log.time_iterate / max(log.iters, 1), 3),
                              "err_last": float(log.errors[-1])}), flush=True)
    if want("srct"):
        # SRCT sketch [S A | S b] at the headline size (pruned two-level DCT on DMMA GEMMs)
        S = rla.srct_operator(d, m, 5)
        W = torch.zeros(d, n + 2, dtype=torch.float64, device="cuda")
        t, tm = timeit(lambda: S.sketch_into(A, u, W[:, :n + 1]), warm=1, reps=2)
        m2 = S.choose_m2(m, d)
        report("srct_sketch", t, tm, bytesA, flops=4.0 * m * (n + 1) * (m2 + d / m2) if m2 else 2.0 * d * m * (n + 1),
               extra={"d": d, "m2": m2, "dense_equiv_flops": 2.0 * d * m * (n + 1)})
        from parla_b200.utils.sketching import SRCTOperator
        SRCTOperator.STAGE_TIMINGS = {}
        S.sketch_into(A, u, W[:, :n + 1])
        print(json.dumps({"srct_stage_seconds": {k: round(v, 4) for k, v in SRCTOperator.STAGE_TIMINGS.items()}}), flush=True)
        SRCTOperator.STAGE_TIMINGS = None
        del S, W
    if want("sps"):
        # saddle-point system at the headline size: SPS2 (LSQR) and SPS1 (PCG; one fused Gram pass per iteration)
        x0 = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        b, _ = K.matvec(A, x0)
        b += 0.1 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
        c = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        delta = 0.5
        for name, alg, tol in (("SPS2-lsqr", rla.SPS2(rla.SkOpSJ(8), 4), 1e-12), ("SPS1-pcg", rla.SPS1(rla.SkOpSJ(8), 4), 1e-12)):
            for rep in range(2):
                x, y, log = alg(A, b, c, delta, tol, 100, 3, logging=True)
            g1 = K.rmatvec(A, y)[:n] - delta * x - c                 # A'y - delta x - c  (second block row)
            tot = log.time_sketch + log.time_factor + log.time_convert + log.time_presolve + log.time_iterate
            print(json.dumps({"solve": name, "m": m, "n": n, "delta": delta, "total_s": round(tot, 4),
                              "sketch": round(log.time_sketch, 4), "factor": round(log.time_factor, 4),
                              "convert": round(log.time_convert, 4), "presolve": round(log.time_presolve, 4),
                              "iterate": round(log.time_iterate, 4), "iters": log.iters,
                              "ms_per_iter": round(1e3 * log.time_iterate / max(log.iters, 1), 3),
                              "block2_resid_rel": float(torch.linalg.vector_norm(g1) / torch.linalg.vector_norm(c)),
                              "err_first_last": [float(log.errors[0]), float(log.errors[-1])]}), flush=True)


if __name__ == "__main__":
    main()
