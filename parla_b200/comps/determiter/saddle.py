"""Preconditioned saddle-system solvers on the device.

Mirrors the interface parla/comps/determiter/saddle.py:20-85, the LSQR-backed implementation
PcSS2 (:180-217) and the PCG-backed PcSS1 (:88-176).  The over-determined branch of PcSS2
(:193-201) is the hot path; its under-determined branch (:203-214) and PcSS1 are "next" rows of
SURVEY.md 8(f).
"""
import torch

from ... import kernels as K
from ...parallel import allreduce_, peer_comm, unwrap
from .lsqr import lsqr, lsqr_adjoint
from .pcg import pcg
from ..preconditioning import a_lift_precond

F64 = torch.float64


def pcss1(A, b, c, delta, tol, iter_lim, R, upper_tri, z0):
    """saddle.py:8-11."""
    return PcSS1()(A, b, c, delta, tol, iter_lim, R, upper_tri, z0)


def pcss2(A, b, c, delta, tol, iter_lim, R, upper_tri, z0):
    """saddle.py:14-17."""
    return PcSS2()(A, b, c, delta, tol, iter_lim, R, upper_tri, z0)


class PrecondSaddleSolver:

    def __call__(self, A, b, c, delta, tol, iter_lim, R, upper_tri, z0):
        raise NotImplementedError()

    exec = __call__


class GramOperator:
    """vec -> A^T (A vec), summed over row shards, in ONE read of A (the reference's ``mv_gram``,
    saddle.py:144-148, reads A twice).  ``delta * vec`` is added by the PCG kernels."""

    def __init__(self, A):
        self.A, self.row_offset, self.group = unwrap(A)
        if self.A.stride(1) != 1 and self.A.shape[1] > 1:
            self.A = self.A.contiguous()
        self.m_local, self.n = self.A.shape
        self.u = torch.empty(self.m_local, dtype=F64, device=self.A.device)     # A vec (scratch, write-only)
        self.zss = torch.empty(self.n + 1, dtype=F64, device=self.A.device)
        self.passes = 0
        self.comm = peer_comm(self.group, self.A.device) if self.A.is_cuda else None   # see PrecondOperator

    def _pass(self, **kw):
        zss = K.stream_pass(self.A, comm=self.comm, **kw)
        self.passes += 1
        if self.comm is None:
            allreduce_(zss, self.group)
        return zss

    def __call__(self, vec, istop=None):
        return self._pass(w=vec, u=self.u, sa=1.0, su=0.0, zss=self.zss, flags=K.PASS_DOT | K.PASS_AXPY, istop=istop)

    def rmatvec(self, y):
        """A^T y summed over row shards (fresh tensor)."""
        return self._pass(u=y, flags=K.PASS_AXPY)[:self.n].clone()

    def residual(self, x, b):
        """b - A x on this rank's rows."""
        y = b.clone()
        K.stream_pass(self.A, w=x, u=y, sa=-1.0, su=1.0, flags=K.PASS_DOT)
        self.passes += 1
        return y


class PcSS1(PrecondSaddleSolver):
    """PCG on the normal equations (A'A + delta I) x = A'b - c, preconditioned by M M' with M = R
    (full rank) or by R~ R~' + (I - V V') for a low-rank SVD-type R (saddle.py:96-176)."""

    ERROR_METRIC_INFO = """
        2-norm of the residual from the normal equations
    """

    def __init__(self):
        self.last_op = None

    def __call__(self, A, b, c, delta, tol, iter_lim, R, upper_tri, z0, _rhs=None):
        m, n = A.shape
        A_loc, _, group = unwrap(A)
        dev = A_loc.device
        b_loc = None if b is None else unwrap(b)[0]
        if b_loc is None:
            b_loc = torch.zeros(A_loc.shape[0], dtype=F64, device=dev)
        if b_loc.dim() != 1:
            raise NotImplementedError()
        if upper_tri:
            raise NotImplementedError()                                       # :117-118
        gram = GramOperator(A)
        self.last_op = gram
        pc_dim = R.shape[1]
        fullrank = pc_dim == n
        if fullrank:
            Rm, gain = R.contiguous(), None
        else:
            # R = V / s  ->  V (columns normalised), R~ = V / t with t = sqrt(s^2 + delta) / its last entry
            # (:122-131);  R~ R~' + I - V V' = I + V diag(1/t^2 - 1) V'   (:134-142 in one product pair)
            sv = 1.0 / torch.linalg.vector_norm(R, dim=0)
            Rm = (R * sv).contiguous()
            t = torch.sqrt(sv * sv + delta)
            t = t / t[-1]
            gain = 1.0 / (t * t) - 1.0
        s_buf = torch.empty(n, dtype=F64, device=dev)
        w_buf = torch.empty(pc_dim, dtype=F64, device=dev)

        def mv_pre(vec, istop):
            zss = K.stream_pass(Rm, u=vec, flags=K.PASS_AXPY, istop=istop)     # Rm' vec
            w_buf.copy_(zss[:pc_dim])
            if gain is not None:
                w_buf.mul_(gain)
            K.stream_pass(Rm, w=w_buf, u=s_buf, sa=1.0, su=0.0, flags=K.PASS_DOT, istop=istop)   # Rm w
            if gain is not None:
                s_buf.add_(vec)
            return s_buf

        if _rhs is not None:
            rhs = _rhs
        else:
            rhs = gram.rmatvec(b_loc)                                          # :150-152
            if c is not None:
                rhs = rhs - c
        x0 = None
        if z0 is not None and fullrank:                                        # :154-158
            x0 = torch.zeros(n, dtype=F64, device=dev)
            K.stream_pass(Rm, w=z0, u=x0, sa=1.0, su=0.0, flags=K.PASS_DOT)
        x, residuals = pcg(lambda vec, istop: gram(vec, istop), rhs, mv_pre, iter_lim, tol, x0, delta=delta)
        y = gram.residual(x, b_loc)                                            # :162
        return x, y, residuals

    exec = __call__


class PcSS2(PrecondSaddleSolver):

    ERROR_METRIC_INFO = """
        2-norm of the residual from the preconditioned normal equations
        (preconditioning on the left and right).
    """

    def __init__(self):
        self.last_op = None

    def __call__(self, A, b, c, delta, tol, iter_lim, R, upper_tri, z0, _op=None, _warm=None, _need_y=True,
                 _b_ridge=None):
        k = 1 if (b is None or b.ndim == 1) else b.shape[1]
        A_pc = _op if _op is not None else a_lift_precond(A, delta, R, upper_tri, k)[0]
        self.last_op = A_pc
        if c is None or float(torch.linalg.vector_norm(c)) == 0:
            b_loc = getattr(b, "local", b)
            result = lsqr(A_pc, b_loc, atol=tol, btol=tol, iter_lim=iter_lim, x0=z0, _warm=_warm, _b_ridge=_b_ridge)
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
