"""Preconditioned saddle-system solvers on the device.

Mirrors the interface parla/comps/determiter/saddle.py:20-85 and the LSQR-backed implementation
PcSS2 (:180-217).  Only the over-determined branch (:193-201) is on the hot path; the
under-determined branch (:203-214) is a "next" row of SURVEY.md 8(f).
"""
import torch

from .lsqr import lsqr, lsqr_adjoint
from ..preconditioning import a_lift_precond


class PrecondSaddleSolver:

    def __call__(self, A, b, c, delta, tol, iter_lim, R, upper_tri, z0):
        raise NotImplementedError()

    exec = __call__


class PcSS2(PrecondSaddleSolver):

    ERROR_METRIC_INFO = """
        2-norm of the residual from the preconditioned normal equations
        (preconditioning on the left and right).
    """

    def __init__(self):
        self.last_op = None

    def __call__(self, A, b, c, delta, tol, iter_lim, R, upper_tri, z0, _op=None, _warm=None, _need_y=True):
        k = 1 if (b is None or b.ndim == 1) else b.shape[1]
        A_pc = _op if _op is not None else a_lift_precond(A, delta, R, upper_tri, k)[0]
        self.last_op = A_pc
        if c is None or float(torch.linalg.vector_norm(c)) == 0:
            b_loc = getattr(b, "local", b)
            result = lsqr(A_pc, b_loc, atol=tol, btol=tol, iter_lim=iter_lim, x0=z0, _warm=_warm)
            x = A_pc.precond(result[0])
            # y = b - A x (saddle.py:199) costs a pass over A; SPO discards it, so it may opt out
            y = A_pc.residual_and_atb(x, b_loc) if _need_y else None
            return x, y, result[7]
        if b is None or float(torch.linalg.vector_norm(getattr(b, "local", b))) == 0:
            # Under-determined least squares (saddle.py:203-214): LSQR on A_pc^T with rhs M^T c
            c_pc = A_pc.precond_t(c)
            y, _y_ridge, _istop, _itn, arnorms, _ = lsqr_adjoint(A_pc, c_pc, atol=tol, btol=tol, iter_lim=iter_lim)
            n = A_pc.n
            if delta > 0:
                x = (A_pc.rmatvec_plain(y) - c) / delta
            else:
                x = torch.full((n,), float("nan"), dtype=y.dtype, device=y.device)
            return x, y, arnorms
        raise ValueError('One of "b" or "c" must be zero.')

    exec = __call__
