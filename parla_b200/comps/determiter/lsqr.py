"""LSQR with every vector and scalar resident on the device.

Mirrors parla/comps/determiter/lsqr.py:98-574 (Paige-Saunders LSQR as modified by PARLA: arnorm
history :413/:572, direct use of x0 :367-370, early return :392-395, stopping rules :500-526).
The host loop below only ENQUEUES work: per iteration one triangular solve, one fused pass over
A, (an all-reduce when row-sharded,) one transposed solve and one recurrence kernel.  The
convergence flag lives in device memory; kernels turn into no-ops once it is set, so the host
polls it a couple of iterations late through pinned memory instead of synchronising every step.
"""
import os

import numpy as np
import torch

from ... import kernels as K
from ...parallel import allreduce_

F64 = torch.float64
POLL_LAG = 2      # iterations the host may run ahead of the device-side stopping test
_POLL_CACHE = {}
# Optional (PLA_LSQR_GRAPH=1): capture one iteration into a CUDA graph and replay it (the device-side stop flag
# already turns every kernel of a replay after convergence into a no-op; single GPU only, no collective inside).
# OFF by default: measured on B200 at 2^16 x 500 (36 iterations) the eager loop takes 0.11 ms per iteration and
# the replayed graph 0.22 ms -- the iteration is bound by the dependent chain of 8 short kernels, not by the host,
# which already runs POLL_LAG iterations ahead (profiles/r2_cfg1.jsonl).
GRAPH_MAX_ELEMS = int(os.environ.get("PLA_LSQR_GRAPH_MAX_ELEMS", 1 << 27))
USE_GRAPH = os.environ.get("PLA_LSQR_GRAPH", "0") == "1"
_CAPTURE_STREAMS = {}
# Fused vector phase (PLA_LSQR_FUSED, default on): for a small dense preconditioner on one GPU the six launches
# between two passes over A (reduce, M^T z + its reduce + copy, recurrences, M v) become one cluster kernel,
# K.lsqr_fused_step -- at 2^16 x 500 they cost more than the pass itself (profiles/r2_cfg1_final.jsonl).
FUSED_MAX_ELEMS = int(os.environ.get("PLA_LSQR_FUSED_MAX_ELEMS", 1 << 20))
USE_FUSED = os.environ.get("PLA_LSQR_FUSED", "1") != "0"


def _capture_iteration(dev, n_max, body):
    """Capture ``body()`` (kernel launches on the current stream, no allocation) into a CUDA graph."""
    key = torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream()
    side, cur = _CAPTURE_STREAMS[key], torch.cuda.current_stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        K.reserve_pass_workspace(dev, n_max)          # per-stream scratch exists before the capture starts
        side.wait_stream(cur)
        l0 = K.launch_count()
        graph.capture_begin(capture_error_mode="thread_local")
        body()
        graph.capture_end()
        nodes = K.launch_count() - l0
    return graph, nodes


def _poll_buffers():
    """Pinned mirrors of the device-side LSQR flags + their events (allocated once per process:
    cudaHostAlloc is far slower than an LSQR iteration on small problems)."""
    key = torch.cuda.current_device()
    if key not in _POLL_CACHE:
        _POLL_CACHE[key] = ([torch.zeros(K.LSQR_NINT, dtype=torch.int32).pin_memory() for _ in range(POLL_LAG + 1)],
                            [torch.cuda.Event() for _ in range(POLL_LAG + 1)])
    return _POLL_CACHE[key]


def lsqr(A, b, damp=0.0, atol=1e-8, btol=1e-8, conlim=1e8, iter_lim=None, show=False, calc_var=False,
         x0=None, _warm=None, _b_ridge=None):
    """A: PrecondOperator; b: device vector (this rank's rows).  Returns the reference's 10-tuple
    (x, istop, itn, r1norm, r2norm, anorm, acond, arnorms, xnorm, var); x is a device tensor.
    ``_b_ridge`` (n-vector, only with delta > 0) is the part of the right-hand side that sits on the
    implicit ridge rows ``sqrt(delta) I`` (zero in SPO; SPS2 puts ``-v[d:]`` there, saddlesys.py:293-294)."""
    if damp != 0.0 or calc_var:
        raise NotImplementedError("PARLA always calls lsqr with damp=0, calc_var=False")
    n = A.shape[1]
    dev = b.device
    if iter_lim is None:
        iter_lim = 2 * n
    iter_lim = int(iter_lim)

    x, v, w = (torch.empty(n, dtype=F64, device=dev) for _ in range(3))
    xw = torch.empty(A.n, dtype=F64, device=dev)
    t = torch.empty(n, dtype=F64, device=dev)
    dstate = torch.zeros(K.LSQR_NDOUBLE, dtype=F64, device=dev)
    istate = torch.zeros(K.LSQR_NINT, dtype=torch.int32, device=dev)
    hist = torch.full((iter_lim,), -1.0, dtype=F64, device=dev)
    bsq = allreduce_(K.sumsq(b), A.group)
    if _b_ridge is not None:
        bsq = bsq + K.sumsq(_b_ridge)

    if _warm is not None:                       # presolve already did  u = b - A_pc x0  (lsqr.py:367-370)
        u, ub, zss = _warm["u"], _warm["ub"], _warm["zss"]
        t.copy_(_warm["t"])
    else:
        u = b.clone()
        ub = None
        if A.delta > 0:
            ub = torch.zeros(A.n, dtype=F64, device=dev) if _b_ridge is None else _b_ridge.clone()
        zss = torch.empty(A.n + 1, dtype=F64, device=dev)
        if x0 is None:
            A.adjoint_pass(u, None if _b_ridge is None else ub, zss, t)     # u = b, beta = |b|, A^T b   (:363-366,:372-375)
            A.atb = zss[:A.n].clone()
        else:
            A.bidiag_pass(x0, u, ub, zss, xw, t, sa=-1.0, su=1.0)
    # the recurrence kernels read |u~|^2 at index len(x) of their `zss` argument; zss holds it at index A.n, which
    # is len(x) only for a full-rank preconditioner (rank-truncated SVD mode: len(x) = rank < A.n)
    zs = zss[A.n - n:]
    K.lsqr_init(t, zs, bsq, atol, btol, conlim, iter_lim, x0, x, v, w, dstate, istate)

    sc = dstate[K.LSQR_SA:K.LSQR_SA + 2]
    istop_dev = istate[0:1]
    pinned, events = _poll_buffers()

    def post(slot):
        pinned[slot].copy_(istate, non_blocking=True)
        events[slot].record()

    def iteration():
        A.bidiag_pass(v, u, ub, zss, xw, t, sc=sc, istop=istop_dev)
        K.lsqr_step(t, zs, x, v, w, dstate, istate, hist)

    fused = USE_FUSED and A.fusable(FUSED_MAX_ELEMS)

    fused_step = K.FusedIteration(A.A, A.R, xw, u, sc, istop_dev, zss, t, x, v, w, dstate, istate, hist) if fused else None

    def iteration_fused():              # xw = M v is already there: the previous fused step (or the prologue) left it
        fused_step()
        A.passes += 1

    post(0)
    events[0].synchronize()
    launched = 0
    if int(pinned[0][0]) == 0:
        graph = None
        if fused:
            A.precond(v, out=xw, istop=istop_dev)
            iteration = iteration_fused
        if USE_GRAPH and A.group is None and A.m_local * A.n <= GRAPH_MAX_ELEMS and iter_lim > 2:
            passes0 = A.passes
            graph, nodes = _capture_iteration(dev, max(A.n, n), iteration)
            A.passes = passes0                  # capturing enqueues nothing
        for it in range(iter_lim):
            if it >= POLL_LAG:
                slot = (it - POLL_LAG) % (POLL_LAG + 1)
                events[slot].synchronize()
                if int(pinned[slot][0]) != 0:
                    break
            if graph is None:
                iteration()
            else:
                graph.replay()
                A.passes += 1
                K.note_launches(nodes)
            launched += 1
            post(it % (POLL_LAG + 1))
    torch.cuda.current_stream().synchronize()
    ds = dstate.cpu().numpy()
    ist = istate.cpu().numpy()
    istop, itn = int(ist[0]), int(ist[1])
    A.passes -= max(0, launched - itn)          # run-ahead launches were device-side no-ops
    var = np.zeros(n)
    if istop == 100:                            # alfa*beta == 0: scalar arnorm, as the reference (:392-395)
        return x, 0, 0, ds[14], ds[14], 0.0, 0.0, np.float64(ds[11]), 0.0, var
    arn = hist[:itn].cpu().numpy()
    return x, istop, itn, ds[14], ds[14], ds[4], ds[13], arn, ds[12], var


LSU_WW = 23      # dstate slot holding |w|^2 (see csrc/lsqr_step.cu)


def lsqr_adjoint(A, c_pc, atol=1e-8, btol=1e-8, conlim=1e8, iter_lim=None):
    """LSQR applied to ``A.T`` for a PrecondOperator ``A`` (what saddle.py:206 does with ``A_pc.T``):
    minimum-norm y with (A_pc)^T y = c_pc.  The long vectors (v~, w, y) live on the rows of A (and on
    the n ridge rows when delta > 0); the same fused pass as the over-determined case does
    v~ <- A M u - (beta/alfa) v~ and z = A^T v~ in one read of A.
    Returns (y_top, y_ridge_or_None, istop, itn, arnorms, dstate)."""
    r = A.shape[1]
    dev = c_pc.device
    if iter_lim is None:
        iter_lim = 2 * A.shape[0]
    iter_lim = int(iter_lim)
    nA, m_loc = A.n, A.m_local
    ridge = A.delta > 0
    u = torch.empty(r, dtype=F64, device=dev)
    t = torch.empty(r, dtype=F64, device=dev)
    xw = torch.empty(nA, dtype=F64, device=dev)
    vt = torch.zeros(m_loc, dtype=F64, device=dev)
    y = torch.zeros(m_loc, dtype=F64, device=dev)
    w = torch.zeros(m_loc, dtype=F64, device=dev)
    vb = torch.zeros(nA, dtype=F64, device=dev) if ridge else None
    yb = torch.zeros(nA, dtype=F64, device=dev) if ridge else None
    wb = torch.zeros(nA, dtype=F64, device=dev) if ridge else None
    zss = torch.empty(nA + 1, dtype=F64, device=dev)
    dstate = torch.zeros(K.LSQR_NDOUBLE, dtype=F64, device=dev)
    istate = torch.zeros(K.LSQR_NINT, dtype=torch.int32, device=dev)
    hist = torch.full((iter_lim,), -1.0, dtype=F64, device=dev)
    sc = dstate[K.LSQR_SA:K.LSQR_SA + 2]
    istop_dev = istate[0:1]

    def long_update(itn):
        K.lsqr_under_long(vt, y, w, dstate, istate, itn)
        if A.group is not None:                           # |w|^2 over all row shards
            from ...parallel import allreduce_
            allreduce_(dstate[LSU_WW:LSU_WW + 1], A.group)
        if ridge:
            K.lsqr_under_long(vb, yb, wb, dstate, istate, itn, add_to_ww=True)

    K.lsqr_under_init(c_pc, u, atol, btol, conlim, iter_lim, dstate, istate)
    A.bidiag_pass(u, vt, vb, zss, xw, t, sc=sc)           # v~_0 = A_pc u_0 (su = 0), t = M^T A^T v~_0
    K.lsqr_under_init2(nA, zss, dstate, istate)
    long_update(0)

    pinned, events = _poll_buffers()

    def post(slot):
        pinned[slot].copy_(istate, non_blocking=True)
        events[slot].record()

    post(0)
    events[0].synchronize()
    if int(pinned[0][0]) == 0:
        for it in range(iter_lim):
            if it >= POLL_LAG:
                slot = (it - POLL_LAG) % (POLL_LAG + 1)
                events[slot].synchronize()
                if int(pinned[slot][0]) != 0:
                    break
            K.lsqr_under_head(t, u, dstate, istate)
            A.bidiag_pass(u, vt, vb, zss, xw, t, sc=sc, istop=istop_dev)
            K.lsqr_under_tail(nA, zss, dstate, istate, hist)
            long_update(it + 1)
            post(it % (POLL_LAG + 1))
    torch.cuda.current_stream().synchronize()
    ist = istate.cpu().numpy()
    istop, itn = int(ist[0]), int(ist[1])
    if istop == 100:
        return y, yb, 0, 0, np.float64(float(dstate[11])), dstate
    return y, yb, istop, itn, hist[:itn].cpu().numpy(), dstate
