#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the PARLA hot path on B200.

Metric (BASELINE.json): SAP1 (sketch-and-precondition, QR preconditioner + LSQR) least-squares solve
time on a 2^22 x 2048 fp64 system per GPU, SJLT sketch (k = 8, d = 4n), tol 1e-12.  A "step" is one
complete solve (sketch -> Householder QR -> presolve -> LSQR to tolerance -> residual).

  python bench.py [--gpus N --steps K --warmup W]          our arm   (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference [...]                   CPU arm: the oracle port of the reference's
                                                           numpy/scipy path on the box's host cores
  python bench.py --workload lowrank [...]                 BASELINE.json configs[3]: SVD1(QB1(RF1(RS1))) on a
                                                           2^20 x 2^14 matrix, rank 512, 2 power iterations

N > 1 is WEAK scaling over row shards: every rank holds 2^22 rows (m_global = N * 2^22), the only
collectives are all-reduces of the d x (n+1) sketch and of n+1 doubles per LSQR iteration.  The line
also carries, outside the headline timing:
  "gauss"   (N = 1)  the Gaussian (Philox-fused DMMA) sketch of the same A: TFLOP/s and fraction of the FP64
                     tensor-pipe peak measured live by pla_dmma_probe;
  "strong"  (any N)  STRONG scaling on a fixed global 2^22 x 4096 problem (BASELINE.json configs[4] at the
                     size that fits one GPU) split over the N ranks: SAP1 and SAP2 solve times;
  "parity"  (any N)  a small row-sharded solve with a replayed reference-format SJLT checked against the CPU
                     oracle (checker only, outside every timed region).
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
LR_M, LR_N, LR_K, LR_R = 1 << 20, 1 << 14, 512, 2048          # configs[3]
LR_CPU_M, LR_CPU_N, LR_CPU_K, LR_CPU_R = 1 << 14, 1 << 11, 128, 1024   # the CPU-sized low-rank sample


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lsq", choices=["lsq", "lowrank"])
    # (--rows/--cols rather than only --m/--n: torchrun's own parser treats "--m"/"--n" as ambiguous abbreviations)
    ap.add_argument("--rows", "--m", dest="m", type=int, default=None, help="rows per GPU")
    ap.add_argument("--cols", "--n", dest="n", type=int, default=None)
    ap.add_argument("--rank-k", dest="k", type=int, default=LR_K, help="target rank (low-rank workload)")
    ap.add_argument("--sketch", default="sjlt", choices=["sjlt", "gauss"])
    ap.add_argument("--mode", default="qr", choices=["qr", "svd", "chol"], help="SPO preconditioner: qr = SAP1, svd = SAP2")
    ap.add_argument("--cpu-rows", type=int, default=1 << 18, help="rows of the bounded CPU sample (SURVEY 8d: 2^18)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the gauss / strong / parity sub-records")
    ap.add_argument("--strong-rows", type=int, default=1 << 22, help="global rows of the strong-scaling sub-record")
    a = ap.parse_args()
    if a.m is None:
        a.m = (1 << 22) if a.workload == "lsq" else LR_M
    if a.n is None:
        a.n = 2048 if a.workload == "lsq" else LR_N
    return a


def metric_name(a):
    """BASELINE.json's metric for the default workload; a descriptive name for any other shape / mode."""
    if a.workload == "lowrank":
        return f"svd1_qb1_lowrank_time_{a.m}x{a.n}_k{a.k}_fp64"
    if (a.m, a.n, a.mode) == (1 << 22, 2048, 'qr'):
        return METRIC
    return f"{'sap1' if a.mode == 'qr' else 'sap2' if a.mode == 'svd' else 'spo_chol'}_lsq_solve_time_{a.m}x{a.n}_per_gpu_fp64"


def workload_name(a, world):
    if a.workload == "lowrank":
        return (f"SVD1(QB1(RF1(RS1(SkOpGA, 2 power iterations, orth)))) randomized low-rank, {a.m}x{a.n} fp64 per GPU "
                f"(m_global={a.m * world}), decaying spectrum exp(-i/100) of rank {LR_R}, target rank {a.k} "
                + ("[BASELINE.json configs[3]]" if (a.m, a.n, a.k) == (LR_M, LR_N, LR_K) else "[non-default shape]"))
    return (f"{'SAP1' if a.mode == 'qr' else 'SAP2' if a.mode == 'svd' else 'SPO'}/SPO(mode={a.mode}) overdetermined least squares, {a.m}x{a.n} fp64 per GPU "
            f"(m_global={a.m * world}), {a.sketch.upper()} sketch"
            f"{' k=8' if a.sketch == 'sjlt' else ''}, d=4n={SF * a.n}, tol=1e-12, iter_lim=100 "
            + ("[BASELINE.json configs[1]]" if (a.m, a.n, a.mode) == (1 << 22, 2048, 'qr') else
               "[BASELINE.json configs[4] shape]" if a.n == 4096 else "[non-default shape]"))


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_SAMPLE = {}


def cpu_solve_sample(n, rows, seed=0):
    """One oracle (numpy/scipy port of the reference) SPO solve on a `rows` x n sample of the
    workload; returns the phase times.  All host BLAS threads."""
    import numpy as np
    from oracle import parla_oracle as orc
    if (n, rows) not in _CPU_SAMPLE:
        # one synthetic system per run (generating 2^18 x 2048 normals takes as long as solving it); the steps differ
        # in the sketching operator's seed, like the GPU arm's
        rng = np.random.default_rng(0)
        A = rng.standard_normal((rows, n))
        _CPU_SAMPLE.clear()
        _CPU_SAMPLE[(n, rows)] = (A, A @ rng.standard_normal(n) + 0.1 * rng.standard_normal(rows))
    A, b = _CPU_SAMPLE[(n, rows)]
    x, log = orc.SPO(orc.SkOpSJ(VEC_NNZ), SF, MODE)(A, b, 0.0, TOL, ITER_LIM, np.random.default_rng(seed + 1))
    return dict(sketch=log.time_sketch, factor=log.time_factor, presolve=log.time_presolve,
                iterate=log.time_iterate, iters=int(log.errors.size - 1))


def cpu_extrapolate(ph, rows, m_full):
    """Every phase but the d x n factorisation is O(m): scale those by m_full / rows."""
    f = m_full / rows
    return (ph["sketch"] + ph["presolve"] + ph["iterate"]) * f + ph["factor"]


def cpu_lowrank_sample(seed=0):
    """Oracle SVD1(QB1(RF1(RS1(2)))) on the CPU-sized low-rank sample (2^14 x 2^11, k = 128)."""
    import numpy as np
    from oracle import parla_oracle as orc
    A = orc.exponent_spectrum(LR_CPU_M, LR_CPU_N, LR_CPU_R, np.random.default_rng(seed), 50.0)
    alg = orc.SVD1(orc.QB1(orc.RF1(orc.RS1(orc.SkOpGA(), 2, orc.orth, 1))))
    t0 = time.time()
    alg(A, LR_CPU_K, float('nan'), 0, np.random.default_rng(seed + 1))
    return time.time() - t0


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


def cpu_baseline_record(a, world, ph=None):
    """cpu_baseline of the least-squares workload: measured sample + the stated extrapolation."""
    if ph is None:
        ph = cpu_solve_sample(a.n, a.cpu_rows)
    cores, blas = cpu_threads()
    m_global = a.m * world
    measured = sum(ph[k] for k in ("sketch", "factor", "presolve", "iterate"))
    value = cpu_extrapolate(ph, a.cpu_rows, m_global)
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample_rows": a.cpu_rows, "sample_measured_s": measured, "m_global": m_global,
            "extrapolation_factor_on_O(m)_phases": m_global / a.cpu_rows,
            "sample": (f"oracle port (numpy/scipy, {blas}, {cores} threads) of the reference SPO on a {a.cpu_rows}x{a.n} "
                       f"sample: measured {measured:.2f} s = sketch {ph['sketch']:.2f} + QR {ph['factor']:.2f} + presolve "
                       f"{ph['presolve']:.2f} + LSQR {ph['iterate']:.2f} ({ph['iters']} its). `value` EXTRAPOLATES that to "
                       f"m_global = {m_global} rows: the O(m) phases x{m_global // a.cpu_rows}, the {SF * a.n}x{a.n} QR "
                       f"unscaled. It is not a measurement at full size.")}, ph


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    world = max(1, a.gpus)
    cores, blas = cpu_threads()
    t_all = time.time()
    if a.workload == "lowrank":
        for _ in range(min(a.warmup, 1)):
            cpu_lowrank_sample()
        ts = [cpu_lowrank_sample(seed=i) for i in range(a.steps)]
        t_s = sum(ts) / len(ts)
        f = (a.m * world * a.n * a.k) / (LR_CPU_M * LR_CPU_N * LR_CPU_K)
        v = t_s * f
        sample = (f"oracle port (numpy/scipy, {blas}, {cores} threads) SVD1(QB1(RF1(RS1(2)))) on {LR_CPU_M}x{LR_CPU_N}, "
                  f"k={LR_CPU_K}: measured {t_s:.2f} s per step; `value` EXTRAPOLATES by the flop ratio m n k: x{f:.0f}")
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample_measured_s": t_s,
               "extrapolation_factor": f, "sample": sample}
    else:
        for _ in range(min(a.warmup, 1)):            # (one untimed solve loads the BLAS / page-faults the arrays)
            cpu_solve_sample(a.n, a.cpu_rows)
        vals, ph = [], None
        for i in range(a.steps):
            cpu, ph = cpu_baseline_record(a, world, cpu_solve_sample(a.n, a.cpu_rows, seed=i))
            vals.append(cpu["value"])
        v = sum(vals) / len(vals)
        cpu["value"] = v
    line = {"impl": "reference", "metric": metric_name(a), "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * (time.time() - t_all) / max(a.steps, 1),
            "ms_per_step_is": "wall clock of one bounded CPU sample (see cpu_baseline.sample); `value` is that sample extrapolated to the workload",
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(a, world),
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_of(a, world):
    """Identical for both arms (the driver compares them)."""
    return {"workload": workload_name(a, world), "parallelism": f"row-sharded x{world}",
            "l2": f"inputs ({a.m * a.n * 8 / 2 ** 30:.0f} GiB per GPU) are far larger than the 126 MB L2; no flush needed",
            "timing": "CUDA events around the K steps, max over ranks (reference arm: host wall clock)"}


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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def numa_pin_for_gpu(local):
    """Run this rank's host-side work (and first-touch its pinned buffers) on the NUMA node of its GPU."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def dmma_peak(dev):
    """FP64 tensor-pipe peak measured live (MEASURED_PEAKS.json carries no FP64 figure): pla_dmma_probe runs
    register-resident DMMA.8x8x4 chains, 8 warps x 4 CTAs per SM."""
    import torch
    from parla_b200 import _lib
    lib = _lib.load()
    sink = torch.zeros(1 << 16, dtype=torch.float64, device=dev)
    iters, cps = 20000, 4
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.pla_dmma_probe(iters, cps, sink.data_ptr(), torch.cuda.current_stream().cuda_stream), "probe")
        e1.record()
        torch.cuda.synchronize()
        flops = 2.0 * 8 * 8 * 4 * 8 * iters * 8 * cps * lib.pla_num_sms()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def parity_record(rla, rank, world, dev):
    """CHECKER, outside every timed region: a small problem in numpy, the reference-format SJLT generated by the
    oracle's restatement of parla/utils/sketching.py:34-80 replayed on the device path (row-sharded over the
    ranks), x compared with the CPU oracle's x."""
    import numpy as np
    import torch
    from oracle import parla_oracle as orc
    m, n = 16384, 256
    rng = np.random.default_rng(12345)
    A = rng.standard_normal((m, n)) * np.logspace(0, 2, n)
    b = A @ rng.standard_normal(n) + 0.1 * rng.standard_normal(m)
    S = orc.sjlt_operator(SF * n, m, np.random.default_rng(99), VEC_NNZ)
    replay = lambda r, c, g: S
    rows = m // world
    lo = rank * rows
    Ad = torch.from_numpy(A[lo:lo + rows].copy()).to(dev)
    bd = torch.from_numpy(b[lo:lo + rows].copy()).to(dev)
    if world > 1:
        Ad, bd = rla.RowSharded(Ad, lo, m), rla.RowSharded(bd, lo, m)
    out = {}
    for name, cls, mode in (("sap1", rla.SAP1, 'qr'), ("sap2", rla.SAP2, 'svd')):
        x, log = cls(replay, SF)(Ad, bd, 0.0, TOL, ITER_LIM, None, logging=True)
        if rank == 0:
            x_ref, log_ref = orc.SPO(replay, SF, mode)(A, b, 0.0, TOL, ITER_LIM, None)
            x = x.cpu().numpy()
            out[name] = {"rel_err_vs_oracle": float(np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)),
                         "iters": int(log.iters), "iters_oracle": int(log_ref.errors.size - 1)}
    out["problem"] = f"{m}x{n} cond 1e2, reference-format SJLT replayed, row-sharded x{world}; tolerance 1e-10"
    return out


def strong_record(rla, K, a, rank, world, dev, barrier, dist):
    """STRONG scaling: a fixed global `--strong-rows` x 4096 problem split over the ranks (137 GB at the default
    2^22 rows; BASELINE.json configs[4] is 2^24 x 4096, which needs >= 4 GPUs), SAP1 and SAP2, SJLT d = 4n."""
    import torch
    n = 4096
    mg = a.strong_rows
    rows = mg // world
    free, _ = torch.cuda.mem_get_info()
    if rows * n * 8 * 1.12 + SF * n * (n + 2) * 8 * 3 > free:
        return {"skipped": f"shard of {rows}x{n} does not fit in the {free / 1e9:.0f} GB free on this GPU"}
    g = torch.Generator(device=dev).manual_seed(2000 + rank)
    A = torch.randn(rows, n, dtype=torch.float64, device=dev, generator=g)
    x0 = torch.randn(n, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    b, _ = K.matvec(A, x0)
    b += 0.1 * torch.randn(rows, dtype=torch.float64, device=dev, generator=g)
    Ash = rla.RowSharded(A, rank * rows, mg) if world > 1 else A
    bsh = rla.RowSharded(b, rank * rows, mg) if world > 1 else b
    rec = {"m_global": mg, "n": n, "rows_per_gpu": rows, "sketch": "sjlt k=8 d=16384", "tol": TOL}
    for name, cls in (("sap1", rla.SAP1), ("sap2", rla.SAP2)):
        alg = cls(rla.SkOpSJ(VEC_NNZ), SF)
        alg(Ash, bsh, 0.0, TOL, ITER_LIM, 1, logging=False)                      # warm-up
        x, log = alg(Ash, bsh, 0.0, TOL, ITER_LIM, 2, logging=True)              # phase breakdown (untimed)
        reps = 2
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            x, _ = alg(Ash, bsh, 0.0, TOL, ITER_LIM, 3 + i, logging=False)
        e1.record()
        barrier()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt)
        rec[name + "_s"] = t
        rec[name + "_phases_s"] = dict(sketch=log.time_sketch, factor=log.time_factor, presolve=log.time_presolve,
                                       iterate=log.time_iterate, iters=log.iters)
        rec[name + "_rel_err_vs_x0"] = float(torch.linalg.vector_norm(x - x0) / torch.linalg.vector_norm(x0))
    del A, b, Ash, bsh
    torch.cuda.empty_cache()
    return rec


def run_gpu_arm(a):
    # a freed 64 GiB block must really go back to the driver before the 128 GiB strong-scaling shard is allocated
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
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
    numa = numa_pin_for_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dmma_burst = dmma_peak(dev) if (rank == 0 and not a.no_extras) else None      # idle, cool GPU: burst clocks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if a.workload == "lowrank":
        return run_lowrank(a, rla, K, rank, world, local, dev, barrier, dist)
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
    peaks = load_peaks()
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback")
    alg_bytes = m * n * 8 + 2 * m * 8             # A read once + u read and written
    t_pass = sum(fused) / max(len(fused), 1)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "stream_pass_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"kernel": "pla::stream_pass_kernel<2,4,2,256> (fused A w / A^T u, one read of A per LSQR iteration)",
                "bound": "hbm", "achieved": alg_bytes / t_pass / 1e9 if fused else None, "peak": peak, "unit": "GB/s",
                "frac": (alg_bytes / t_pass / 1e9 / peak) if fused else None, "traffic": traffic,
                "peak_source": peak_src, "launches_timed": len(fused), "ms_per_launch": 1e3 * t_pass,
                "share_of_step": (sum(fused) / a.steps) / t_step if fused else None}
    if fused and roofline["frac"] > 1.0:
        roofline["note"] = ("above 1: the peak is the driver's COPY figure (reads + writes share the bus); this kernel "
                            "only reads A, and read-only streams measure up to ~7 % faster than copies on these boxes")

    # residual check of the last timed solve (result is used, nothing is skipped)
    r, zss = K.matvec(A, x, alpha=1.0, y=b.clone(), beta=-1.0)
    atr = K.rmatvec(A, r)
    if world > 1:
        dist.all_reduce(atr)
    check = {"rel_normal_eq_residual": float(torch.linalg.vector_norm(atr[:n]) / (math.sqrt(float(atr[n])) * math.sqrt(m * world))),
             "rel_err_vs_x0": float(torch.linalg.vector_norm(x - x0) / torch.linalg.vector_norm(x0))}
    del r

    # ---- sub-record: the Gaussian (Philox-fused DMMA) sketch of the same A (configs[1] "Gaussian vs SJLT")
    extras = {}
    if not a.no_extras and world == 1:
        d = SF * n
        pk = dmma_peak(dev)
        op = rla.gaussian_operator(d, m, 11)
        W = torch.zeros(d, n + 2, dtype=torch.float64, device=dev)
        op = rla.as_device_operator(op, dev)
        op.sketch_into(A, b, W[:, :n + 1])                   # warm-up (one full sketch)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); op.sketch_into(A, b, W[:, :n + 1]); g1.record()
        torch.cuda.synchronize()
        tg = g0.elapsed_time(g1) * 1e-3
        flops = 2.0 * d * m * (n + 1)
        # sketch == the same operator applied as a dense matrix on a row/column sample (fp32 Box-Muller inside)
        extras["gauss"] = {"kernel": "pla::gauss_sketch_kernel (S tiles generated in shared memory from Philox4x32-10, DMMA.8x8x4)",
                           "sketch_s": tg, "flops": flops, "achieved": flops / tg / 1e12, "unit": "TFLOP/s",
                           "peak": pk, "frac": flops / tg / 1e12 / pk, "bound": "tensor (FP64 DMMA)",
                           "peak_source": "pla_dmma_probe measured live right before the sketch, i.e. on the hot, power-capped GPU "
                                          "(MEASURED_PEAKS.json has no FP64 figure); peak_burst is the same probe on the idle GPU at "
                                          "the start of the run",
                           "peak_burst": dmma_burst, "frac_of_burst": flops / tg / 1e12 / dmma_burst if dmma_burst else None,
                           "sjlt_sketch_s_same_A": phases["sketch"]}
        del W, op

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
        del A, b, Ash, bsh
        torch.cuda.empty_cache()
        Ahs = rla.RowSharded(Ah, rank * rows_host, world * rows_host) if world > 1 else Ah
        bhs = rla.RowSharded(bh, rank * rows_host, world * rows_host) if world > 1 else bh
        e2e_steps = max(1, min(a.steps, 3))

        def e2e_solve(seed):
            xx, _ = alg(Ahs, bhs, 0.0, TOL, ITER_LIM, seed, logging=False)     # host (pinned) shards in, host x out
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
        up = getattr(alg, "last_upload", None)
        e2e = {"value": t_e2e, "unit": UNIT, "h2d_bytes_per_step": rows_host * n * 8 + rows_host * 8,
               "d2h_bytes_per_step": n * 8, "steps": e2e_steps, "host_buffers": note, "numa_node": numa,
               "h2d_GBps_this_rank": (up["bytes"] / up["seconds"] / 1e9) if up else None,
               "api": "parla_b200.SPO(...)(A_host, b_host, ...) -> x_host  (RowSharded host shards when N > 1)"}
        del Ah, bh, Ahs, bhs
    else:
        del A, b, Ash, bsh
    alg.last_residual = None
    alg.iterative_solver.last_op = None
    x = log = None
    K.Workspace._bufs.clear()
    torch.cuda.empty_cache()

    if not a.no_extras:
        # sub-records outside the headline timing: a failure in one of them must not cost the headline line
        for key, fn in (("strong", lambda: strong_record(rla, K, a, rank, world, dev, barrier, dist)),
                        ("parity", lambda: parity_record(rla, rank, world, dev))):
            try:
                extras[key] = fn()
            except Exception as exc:                                   # noqa: BLE001
                extras[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not a.no_cpu:
        use_all_host_threads()
        cpu, _ = cpu_baseline_record(a, world)

    line = {"metric": metric_name(a), "value": t_step, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_of(a, world),
            "phases_s": phases, "check": check, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "effective_GBps_over_A": phases["passes_over_A"] * m * n * 8 / t_step / 1e9}
    line.update(extras)
    if world > 1:
        from parla_b200 import parallel as par
        fused_x = any(c.ok for c in par._PEER_COMMS.values())
        line["collectives"] = {
            "sketch_sum": "NCCL all-reduce of the d x (n + 1) sketch, once per solve",
            "per_iteration_sum": ("fused into the pass's reduce kernel: one-shot exchange of n + 1 LL lines over NVLink peer "
                                  "memory (pla_stream_pass_peer_f64)") if fused_x else "NCCL all-reduce of n + 1 doubles"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_lowrank(a, rla, K, rank, world, local, dev, barrier, dist):
    """BASELINE.json configs[3]: SVD1(QB1(RF1(RS1(SkOpGA, 2, orth, 1)))) on A = (U sigma) V^T built on the device."""
    import numpy as np
    import torch
    m, n, k, r = a.m, a.n, a.k, min(LR_R, a.n)
    g = torch.Generator(device=dev).manual_seed(rank)
    # A = (U sigma) V^T with orthonormal U (per shard: orthonormal columns on this shard's rows), V replicated
    U = rla.orth(torch.randn(m, r, dtype=torch.float64, device=dev, generator=g))
    V = rla.orth(torch.randn(n, r, dtype=torch.float64, device=dev,
                             generator=torch.Generator(device=dev).manual_seed(12345)))
    sigma = torch.exp(-torch.arange(r, dtype=torch.float64, device=dev) / 100.0)
    U.mul_(sigma)
    if world > 1:
        U.mul_(1.0 / math.sqrt(world))                         # global U has orthonormal columns
    A = K.gemm(U, V, transb=True)
    del U
    torch.cuda.empty_cache()
    Ash = rla.RowSharded(A, rank * m, world * m) if world > 1 else A
    alg = rla.SVD1(rla.QB1(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1))))

    def step(seed):
        return alg(Ash, k, np.nan, 0, seed)

    for i in range(a.warmup):
        step(100 + i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = K.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        Uh, s, Vh = step(i)
    e1.record()
    barrier()
    launches = K.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t_step = e0.elapsed_time(e1) * 1e-3 / a.steps
    if world > 1:
        tt = torch.tensor([t_step], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step = float(tt)
    s_true = sigma[:s.numel()]
    rel = (s - s_true).abs() / s_true
    check = {"max_rel_sv_err_leading_half": float(rel[:max(1, s.numel() // 2)].max()),
             "max_rel_sv_err_all_k": float(rel.max()),
             "note": "no oversampling (over = 0, as configs[3] states): the trailing values of a rank-k QB are under-estimated",
             "orth_V": float(torch.linalg.norm(K.gemm(Vh, Vh, transb=True) - torch.eye(s.numel(), dtype=torch.float64, device=dev)))}
    # roofline of the dominant kernel: the DMMA GEMM of one pass over A (Y = A S: 2 m n k flop), timed alone
    pk = dmma_peak(dev)
    S = torch.randn(n, k, dtype=torch.float64, device=dev, generator=g)
    Y = K.gemm(A, S)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); K.gemm(A, S, out=Y); g1.record(); torch.cuda.synchronize()
        ts.append(g0.elapsed_time(g1) * 1e-3)
    tg = sum(ts) / len(ts)
    flops = 2.0 * m * n * k
    roofline = {"kernel": "pla::gemm_f64_kernel<0,0> (Y = A S, one of the 4 passes over A of SVD1(QB1) with 2 power iterations)",
                "bound": "tensor", "achieved": flops / tg / 1e12, "peak": pk, "unit": "TFLOP/s",
                "frac": flops / tg / 1e12 / pk, "traffic": None, "ms_per_launch": 1e3 * tg,
                "peak_source": "FP64 DMMA peak from pla_dmma_probe measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "share_of_step": 4 * tg / t_step, "flops_per_launch": flops,
                "whole_step_TFLOPs_on_4_passes": 4 * flops / t_step / 1e12}
    del S, Y
    e2e = None
    if not a.no_e2e:
        import psutil
        have = psutil.virtual_memory().available / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        rows_host = m
        note = "full matrix in pinned host memory"
        if have < 1.6 * m * n * 8:
            rows_host = max(1 << 12, int(0.5 * have / (n * 8)) // 4096 * 4096)
            note = (f"host RAM per rank allows {rows_host} pinned rows; the e2e step runs on that many rows and its time is "
                    f"scaled by {m}/{rows_host} (every phase but the {k}x{n} SVD is O(m))")
        Ah = torch.empty(rows_host, n, dtype=torch.float64, pin_memory=True)
        Ah.copy_(A[:rows_host])
        del A, Ash
        torch.cuda.empty_cache()

        def e2e_step(seed):
            Ad = Ah.to(dev, non_blocking=True)
            Ads = rla.RowSharded(Ad, rank * rows_host, world * rows_host) if world > 1 else Ad
            Uh, s, Vh = alg(Ads, k, np.nan, 0, seed)
            return s.cpu(), Vh.cpu()                       # the factors a caller reads back (U stays sharded on the device)
        e2e_step(50)
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(a.steps, 2))
        for i in range(reps):
            e2e_step(i)
        barrier()
        t_e2e = (time.perf_counter() - t0) / reps * (m / rows_host)
        e2e = {"value": t_e2e, "unit": UNIT, "h2d_bytes_per_step": rows_host * n * 8, "d2h_bytes_per_step": (k + k * n) * 8,
               "steps": reps, "host_buffers": note,
               "api": "parla_b200.SVD1(QB1(RF1(RS1(...))))(A_uploaded_inside_the_timing, k, nan, 0, rng) -> (U dev, s host, Vh host)"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not a.no_cpu:
        use_all_host_threads()
        t_s = cpu_lowrank_sample()
        cores, blas = cpu_threads()
        f = (m * n * k) / (LR_CPU_M * LR_CPU_N * LR_CPU_K)
        cpu = {"value": t_s * f, "unit": UNIT, "cores": cores, "kind": "port", "sample_measured_s": t_s,
               "extrapolation_factor": f,
               "sample": (f"oracle port (numpy/scipy, {blas}, {cores} threads) SVD1(QB1(RF1(RS1(2)))) on {LR_CPU_M}x{LR_CPU_N}, "
                          f"k={LR_CPU_K}: measured {t_s:.2f} s; `value` EXTRAPOLATES by the flop ratio m n k (x{f:.0f})")}
    line = {"metric": metric_name(a), "value": t_step, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_of(a, world), "check": check, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
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
