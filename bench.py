#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the PARLA hot path on B200.

Metric (BASELINE.json): SAP1 (sketch-and-precondition, QR preconditioner + LSQR) least-squares solve
time on a 2^22 x 2048 fp64 system per GPU, SJLT sketch (k = 8, d = 4n), tol 1e-12.  A "step" is one
complete solve (sketch -> Householder QR -> presolve -> LSQR to tolerance -> residual).

  python bench.py [--gpus N --steps K --warmup W]          our arm   (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference [...]                   CPU arm: the oracle port of the reference's
                                                           numpy/scipy path on the box's host cores

N > 1 is WEAK scaling over row shards: every rank holds 2^22 rows (m_global = N * 2^22), the only
collectives are all-reduces of the d x (n+1) sketch and of n+1 doubles per LSQR iteration.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL, ITER_LIM, SF, VEC_NNZ = 1e-12, 100, 4, 8
METRIC, UNIT = "sap1_lsq_solve_time_2^22x2048_fp64", "s"
MODE = "qr"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # (--rows/--cols rather than only --m/--n: torchrun's own parser treats "--m"/"--n" as ambiguous abbreviations)
    ap.add_argument("--rows", "--m", dest="m", type=int, default=1 << 22, help="rows per GPU")
    ap.add_argument("--cols", "--n", dest="n", type=int, default=2048)
    ap.add_argument("--sketch", default="sjlt", choices=["sjlt", "gauss"])
    ap.add_argument("--mode", default="qr", choices=["qr", "svd", "chol"], help="SPO preconditioner: qr = SAP1, svd = SAP2")
    ap.add_argument("--cpu-rows", type=int, default=1 << 16, help="rows of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def metric_name(a):
    """BASELINE.json's metric for the default workload; a descriptive name for any other shape / mode."""
    if (a.m, a.n, a.mode) == (1 << 22, 2048, 'qr'):
        return METRIC
    return f"{'sap1' if a.mode == 'qr' else 'sap2' if a.mode == 'svd' else 'spo_chol'}_lsq_solve_time_{a.m}x{a.n}_per_gpu_fp64"


def workload_name(a, world):
    return (f"{'SAP1' if a.mode == 'qr' else 'SAP2' if a.mode == 'svd' else 'SPO'}/SPO(mode={a.mode}) overdetermined least squares, {a.m}x{a.n} fp64 per GPU "
            f"(m_global={a.m * world}), {a.sketch.upper()} sketch"
            f"{' k=8' if a.sketch == 'sjlt' else ''}, d=4n={SF * a.n}, tol=1e-12, iter_lim=100 "
            + ("[BASELINE.json configs[1]]" if (a.m, a.n, a.mode) == (1 << 22, 2048, 'qr') else
               "[BASELINE.json configs[4] shape]" if a.n == 4096 else "[non-default shape]"))


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_solve_sample(n, rows, seed=0):
    """One oracle (numpy/scipy port of the reference) SPO solve on a `rows` x n sample of the
    workload; returns the phase times.  All host BLAS threads."""
    import numpy as np
    from oracle import parla_oracle as orc
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((rows, n))
    b = A @ rng.standard_normal(n) + 0.1 * rng.standard_normal(rows)
    x, log = orc.SPO(orc.SkOpSJ(VEC_NNZ), SF, MODE)(A, b, 0.0, TOL, ITER_LIM, np.random.default_rng(seed + 1))
    return dict(sketch=log.time_sketch, factor=log.time_factor, presolve=log.time_presolve,
                iterate=log.time_iterate, iters=int(log.errors.size - 1))


def cpu_extrapolate(ph, rows, m_full):
    """Every phase but the d x n factorisation is O(m): scale those by m_full / rows."""
    f = m_full / rows
    return (ph["sketch"] + ph["presolve"] + ph["iterate"]) * f + ph["factor"]


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        info = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if info:
            return int(info[0]["num_threads"]), f'{info[0].get("internal_api")} {info[0].get("version")}'
    except Exception:
        pass
    return len(os.sched_getaffinity(0)), "unknown BLAS"


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin the CPU arm to one BLAS thread:
    lift the limit to every core this process may run on (OpenBLAS accepts this at run time)."""
    global _THREAD_LIMITER
    try:
        import numpy, scipy.linalg, scipy.sparse       # noqa: F401,E401  (load the BLAS libraries first)
        from threadpoolctl import threadpool_limits
        _THREAD_LIMITER = threadpool_limits(limits=len(os.sched_getaffinity(0)), user_api="blas")   # keep it alive
    except Exception:
        pass


_THREAD_LIMITER = None


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    world = max(1, a.gpus)
    for _ in range(a.warmup):
        cpu_solve_sample(a.n, a.cpu_rows)
    vals, last = [], None
    t_all = time.time()
    for i in range(a.steps):
        last = cpu_solve_sample(a.n, a.cpu_rows, seed=i)
        vals.append(cpu_extrapolate(last, a.cpu_rows, a.m * world))
    v = sum(vals) / len(vals)
    cores, blas = cpu_threads()
    sample = (f"oracle port (numpy/scipy, {blas}) of the reference SPO on {a.cpu_rows}x{a.n} rows of the workload per "
              f"step (measured {sum(last[k] for k in ('sketch','factor','presolve','iterate')):.2f} s, {last['iters']} "
              f"iterations); sketch/presolve/LSQR phases are O(m) and scaled x{a.m * world // a.cpu_rows}, the "
              f"{SF * a.n}x{a.n} QR is not scaled")
    line = {"impl": "reference", "metric": metric_name(a), "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * (time.time() - t_all) / max(a.steps, 1),
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "timing": "wall clock on host"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_gpu_arm(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import parla_b200 as rla
    from parla_b200 import kernels as K

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m, n = a.m, a.n

    # synthetic problem, generated on the device: b = A x0 + 0.1 noise
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    A = torch.randn(m, n, dtype=torch.float64, device=dev, generator=g)
    x0 = torch.randn(n, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    b, _ = K.matvec(A, x0)
    b += 0.1 * torch.randn(m, dtype=torch.float64, device=dev, generator=g)
    gen = rla.SkOpSJ(VEC_NNZ) if a.sketch == "sjlt" else rla.SkOpGA()
    alg = rla.SPO(gen, SF, a.mode)
    Ash = rla.RowSharded(A, rank * m, world * m) if world > 1 else A
    bsh = rla.RowSharded(b, rank * m, world * m) if world > 1 else b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(seed):
        return alg(Ash, bsh, 0.0, TOL, ITER_LIM, seed, logging=False)

    for i in range(a.warmup):
        solve(100 + i)
    # one logged solve (untimed) for the phase breakdown and the iteration count
    x, log = alg(Ash, bsh, 0.0, TOL, ITER_LIM, 7, logging=True)
    phases = dict(sketch=log.time_sketch, factor=log.time_factor, presolve=log.time_presolve,
                  iterate=log.time_iterate, iters=log.iters, passes_over_A=log.passes_over_A)

    # ---- timed region: EXACTLY K steps, device-timed, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    K.PASS_TIMINGS = []
    l0 = K.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        x, _ = solve(i)
    e1.record()
    barrier()
    launches = K.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t_total = e0.elapsed_time(e1) * 1e-3
    passes = K.PASS_TIMINGS
    K.PASS_TIMINGS = None
    if world > 1:
        tt = torch.tensor([t_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_total = float(tt)
    t_step = t_total / a.steps

    # ---- roofline of the dominant kernel (fused LSQR pass), durations measured live by CUDA events
    fused = [e_a.elapsed_time(e_b) * 1e-3 for fl, mm, nn, e_a, e_b in passes if fl == 3 and mm == m and nn == n]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback")
    alg_bytes = m * n * 8 + 2 * m * 8             # A read once + u read and written
    t_pass = sum(fused) / max(len(fused), 1)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "stream_pass_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": "pla::stream_pass_kernel<2,4,2> (fused A w / A^T u, one read of A per LSQR iteration)",
                "bound": "hbm", "achieved": alg_bytes / t_pass / 1e9 if fused else None, "peak": peak, "unit": "GB/s",
                "frac": (alg_bytes / t_pass / 1e9 / peak) if fused else None, "traffic": traffic,
                "peak_source": peak_src, "launches_timed": len(fused), "ms_per_launch": 1e3 * t_pass,
                "share_of_step": (sum(fused) / a.steps) / t_step if fused else None}

    # residual check of the last timed solve (result is used, nothing is skipped)
    r, zss = K.matvec(A, x, alpha=1.0, y=b.clone(), beta=-1.0)
    atr = K.rmatvec(A, r)
    if world > 1:
        dist.all_reduce(atr)
    check = {"rel_normal_eq_residual": float(torch.linalg.vector_norm(atr[:n]) / (math.sqrt(float(atr[n])) * math.sqrt(m * world))),
             "rel_err_vs_x0": float(torch.linalg.vector_norm(x - x0) / torch.linalg.vector_norm(x0))}

    # ---- e2e: the same solve through the public API from HOST buffers (pinned), H2D/D2H inside the timing
    e2e = None
    if not a.no_e2e:
        import psutil
        need = m * n * 8 + m * 8
        have = psutil.virtual_memory().available / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        note = "full shard in pinned host memory"
        rows_host = m
        if have < 1.6 * need:
            rows_host = max(1 << 14, int(0.6 * have / (n * 8)) // 4096 * 4096)
            note = (f"host RAM per rank allows {rows_host} pinned rows; the e2e solve is run on that many rows and the "
                    f"O(m) part of its time scaled to {m}")
        Ah = torch.empty(rows_host, n, dtype=torch.float64, pin_memory=True)
        bh = torch.empty(rows_host, dtype=torch.float64, pin_memory=True)
        Ah.copy_(A[:rows_host]); bh.copy_(b[:rows_host])
        del A, b, Ash, bsh, r
        torch.cuda.empty_cache()
        Ahs = rla.RowSharded(Ah, rank * rows_host, world * rows_host) if world > 1 else Ah
        bhs = rla.RowSharded(bh, rank * rows_host, world * rows_host) if world > 1 else bh
        e2e_steps = max(1, min(a.steps, 3))
        def e2e_solve(seed):
            if world > 1:      # shards are uploaded by the caller-side helper, still inside the timed region
                Ad, bd = Ah.to(dev, non_blocking=True), bh.to(dev, non_blocking=True)
                xx, _ = alg(rla.RowSharded(Ad, rank * rows_host, world * rows_host),
                            rla.RowSharded(bd, rank * rows_host, world * rows_host), 0.0, TOL, ITER_LIM, seed, logging=False)
                return xx.cpu()
            xx, _ = alg(Ah, bh, 0.0, TOL, ITER_LIM, seed, logging=False)
            return xx
        e2e_solve(50)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            xh = e2e_solve(i)
        barrier()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            tt = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt)
        if rows_host != m:
            t_e2e = t_e2e * (m / rows_host)
        e2e = {"value": t_e2e, "unit": UNIT, "h2d_bytes_per_step": rows_host * n * 8 + rows_host * 8,
               "d2h_bytes_per_step": n * 8, "steps": e2e_steps, "host_buffers": note,
               "api": "parla_b200.SPO(...)(A_host, b_host, ...) -> x_host"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not a.no_cpu:
        use_all_host_threads()
        ph = cpu_solve_sample(n, a.cpu_rows)
        cores, blas = cpu_threads()
        cpu = {"value": cpu_extrapolate(ph, a.cpu_rows, m), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"oracle port (numpy/scipy, {blas}) on {a.cpu_rows}x{n} rows: sketch {ph['sketch']:.2f} s, "
                          f"QR {ph['factor']:.2f} s, presolve {ph['presolve']:.2f} s, LSQR {ph['iterate']:.2f} s "
                          f"({ph['iters']} its); O(m) phases scaled x{m // a.cpu_rows}, QR unscaled")}

    line = {"metric": metric_name(a), "value": t_step, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "parallelism": f"row-sharded x{world}",
                       "l2": f"inputs ({a.m * a.n * 8 / 2 ** 30:.0f} GiB per GPU) are far larger than the 126 MB L2; no flush needed",
                       "timing": "CUDA events around the K solves, max over ranks"},
            "phases_s": phases, "check": check, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "effective_GBps_over_A": phases["passes_over_A"] * m * n * 8 / t_step / 1e9}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    MODE = args.mode
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
