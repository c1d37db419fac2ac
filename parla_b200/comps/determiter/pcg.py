"""Preconditioned conjugate gradients with every vector and scalar resident on the device.

Mirrors parla/comps/determiter/pcg.py:5-47: same loop (history entry i is ||rhs - mat x_i|| at the
start of iteration i, residual recomputed from x on iterations 0, 10, 20, ..., stop when
||r|| <= tol * ||r_0|| or after iter_lim steps).  ``mv_mat`` / ``mv_pre`` are callables that ENQUEUE
work and write into a given buffer; the host loop never reads a device scalar inside an iteration --
the stop flag lives in device memory (kernels become no-ops once it is set) and is polled a couple
of iterations late through pinned memory, exactly as in lsqr.py.
"""
import torch

from ... import kernels as K
from .lsqr import POLL_LAG, _poll_buffers

F64 = torch.float64


def pcg(mv_mat, rhs, mv_pre, iter_lim, tol, x0, delta=0.0):
    """Solve ``(mat + delta I) x = rhs``.

    mv_mat(vec, istop) -> device n-vector ``mat @ vec`` (may alias an internal buffer that the next
    call overwrites); mv_pre(vec, istop) -> ``M M^T vec``.  ``x0`` None means the zero vector (the
    reference's ``rhs - mv_mat(x0)`` is then just ``rhs``).  Returns (x, residuals) with residuals a
    host numpy array of length = iterations taken.
    """
    n = rhs.numel()
    dev = rhs.device
    iter_lim = int(iter_lim)
    x = torch.zeros(n, dtype=F64, device=dev) if x0 is None else x0.clone()
    r = torch.empty(n, dtype=F64, device=dev)
    p = torch.empty(n, dtype=F64, device=dev)
    dstate = torch.zeros(K.LSQR_NDOUBLE, dtype=F64, device=dev)
    istate = torch.zeros(K.LSQR_NINT, dtype=torch.int32, device=dev)
    hist = torch.full((max(iter_lim, 1),), -1.0, dtype=F64, device=dev)
    istop_dev = istate[0:1]

    gx = None if x0 is None else mv_mat(x, None)                     # pcg.py:16
    K.pcg_residual(rhs, gx, delta, x, r, dstate, istate, init=True, tol=tol)
    K.pcg_direction(r, mv_pre(r, None), p, dstate, istate, init=True, iter_lim=iter_lim)   # :18-23

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
            recompute = it % 10 == 0                                  # :34
            K.pcg_update(mv_mat(p, istop_dev), delta, p, x, r, dstate, istate, hist, recompute)   # :28-33,:37
            if recompute:
                K.pcg_residual(rhs, mv_mat(x, istop_dev), delta, x, r, dstate, istate)           # :35
            K.pcg_direction(r, mv_pre(r, istop_dev), p, dstate, istate)                         # :38-43
            post(it % (POLL_LAG + 1))
    torch.cuda.current_stream().synchronize()
    itn = int(istate.cpu()[1])
    return x, hist[:itn].cpu().numpy()
