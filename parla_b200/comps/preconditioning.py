"""Right-preconditioned, ridge-lifted operator on the device.

Mirrors parla/comps/preconditioning.py:16-67 (``a_lift_precond``) and :70-79 (``svd_right_precond``).
Differences from the reference, all behaviour-preserving:
  * ``[A; sqrt(delta) I]`` is never materialised (the reference copies A, :6-13); the identity rows
    are applied by pla_lsqr_ridge_f64 on n-vectors;
  * forward and adjoint products of one LSQR iteration share ONE pass over A
    (:func:`PrecondOperator.bidiag_pass`), see parla_b200/csrc/stream_pass.cu.
"""
import math
import os

import torch

from .. import kernels as K
from ..parallel import allreduce_, peer_comm, unwrap

F64 = torch.float64


class PrecondOperator:
    """A_pc = [A; sqrt(delta) I] M with M = R^{-1} (upper_tri) or M = R (dense n x r)."""

    def __init__(self, A, delta, R, upper_tri):
        self.A, self.row_offset, self.group = unwrap(A)
        if self.A.dim() != 2:
            raise ValueError("A must be 2-D")
        if self.A.stride(1) != 1 and self.A.shape[1] > 1:
            self.A = self.A.contiguous()
        self.m_local, self.n = self.A.shape
        self.m = A.shape[0]
        self.delta = float(delta)
        self.sd = math.sqrt(self.delta)
        self.R = R
        self.tri = bool(upper_tri)
        self.rank = R.shape[1]
        self.shape = (self.m + (self.n if self.delta > 0 else 0), self.rank)
        self.passes = 0
        self.atb = None          # cache of A^T b when a pass happened to produce it
        self._zt = None          # scratch of precond_t (rank + 1 doubles), allocated once
        # row-sharded A on the GPUs of one node: the sum of [z | |u|^2] over the ranks rides inside the pass's reduce
        # kernel (NVLink peer memory); otherwise (None) an NCCL all-reduce follows the pass
        self.comm = peer_comm(self.group, self.A.device) if self.A.is_cuda else None

    def fusable(self, max_elems):
        """True when K.lsqr_fused_step can stand in for reduce -> M^T z -> step -> M v: one GPU, no ridge rows, a small
        dense row-major M, and an A the plain (not column-blocked) streaming pass handles."""
        n, r = self.n, self.rank
        return (self.group is None and self.comm is None and self.delta == 0 and not self.tri
                and self.R.dim() == 2 and self.R.shape == (n, r) and self.R.stride(1) == 1 and self.R.stride(0) >= r
                and n * r <= max_elems and r <= K.FUSED_MAX_R and n <= K.FUSED_MAX_NIN
                and n <= K.PASS_MAX_N and not (n % 2 == 1 and n > K.PASS_MAX_N // 2))

    def _pass(self, **kw):
        """One streaming pass over this rank's rows; ``zss`` comes back summed over the ranks."""
        zss = K.stream_pass(self.A, comm=self.comm, **kw)
        self.passes += 1
        if self.comm is None:
            allreduce_(zss, self.group)
        return zss

    # ---- M and M^T on n-vectors (preconditioning.py:40-41 / :56-57)
    def precond(self, z, out=None, istop=None):
        if self.tri:
            return K.trsv_upper(self.R, z, trans=False, out=out, istop=istop)
        if out is None:
            out = torch.zeros(self.n, dtype=F64, device=z.device)
        K.stream_pass(self.R, w=z, u=out, sa=1.0, su=0.0, flags=K.PASS_DOT, istop=istop)
        return out

    def precond_t(self, w, out=None, istop=None):
        if self.tri:
            return K.trsv_upper(self.R, w, trans=True, out=out, istop=istop)
        if out is None:
            return K.stream_pass(self.R, u=w, flags=K.PASS_AXPY, istop=istop)[:self.rank]
        if self._zt is None:
            self._zt = torch.empty(self.rank + 1, dtype=F64, device=w.device)
        K.stream_pass(self.R, u=w, zss=self._zt, flags=K.PASS_AXPY, istop=istop)
        out.copy_(self._zt[:self.rank])     # (after LSQR has stopped nothing downstream reads `out`)
        return out

    # ---- one Golub-Kahan half-step pair in a single read of A
    def bidiag_pass(self, z, u, ub, zss, xw, t, sc=None, sa=1.0, su=0.0, istop=None):
        """u~ <- sa * A_pc z + su * u~ (top part in ``u``, ridge part in ``ub``);
        t <- M^T [A; sd I]^T u~ ;  zss[n] <- |u~|^2.   (preconditioning.py:26-38 forward + adjoint)"""
        self.precond(z, out=xw, istop=istop)
        self._pass(w=xw, u=u, sc=sc, sa=sa, su=su, zss=zss, flags=K.PASS_DOT | K.PASS_AXPY, istop=istop)
        if self.delta > 0:
            K.lsqr_ridge(self.sd, xw, ub, zss, sc=sc, sa=sa, su=su, istop=istop)
        self.precond_t(zss[:self.n], out=t, istop=istop)
        return t

    def adjoint_pass(self, u, ub, zss, t):
        """t <- M^T (A^T u + sd * ub), zss[n] <- |[u; ub]|^2 (no forward product)."""
        self._pass(u=u, zss=zss, flags=K.PASS_AXPY)
        if self.delta > 0 and ub is not None:
            K.lsqr_ridge(self.sd, None, ub, zss, sa=0.0, su=1.0)
        self.precond_t(zss[:self.n], out=t)
        return t

    def rmatvec_plain(self, y):
        """A^T y (no preconditioner, no ridge rows), summed over row shards."""
        zss = self._pass(u=y, flags=K.PASS_AXPY)
        return zss[:self.n].clone()

    def matvec_plain(self, x):
        """A x on this rank's rows."""
        y = torch.zeros(self.m_local, dtype=F64, device=x.device)
        K.stream_pass(self.A, w=x, u=y, sa=1.0, su=0.0, flags=K.PASS_DOT)
        self.passes += 1
        return y

    def residual_and_atb(self, x, b):
        """y = b - A x and A^T b in one pass (saddle.py:199 and least_squares.py:361 fused)."""
        y = b.clone()
        zss = self._pass(w=x, u=y, g=b, sa=-1.0, su=1.0, flags=K.PASS_DOT | K.PASS_AXPY | K.PASS_AXPY_G)
        self.atb = zss[:self.n].clone()
        return y


def a_lift(A, scale):
    """parla/comps/preconditioning.py:6-13: ``[A; scale * I]`` (A itself when scale == 0).  Provided for API
    completeness -- it COPIES A; nothing on the solver path calls it (the ridge rows stay implicit)."""
    if scale == 0:
        return A
    A = unwrap(A)[0]
    n = A.shape[1]
    return torch.cat((A, scale * torch.eye(n, dtype=A.dtype, device=A.device)), dim=0)


def a_lift_precond(A, delta, R, upper_tri=False, k=1):
    """Signature of parla/comps/preconditioning.py:16.  Returns (A_precond, M_fwd, M_adj)."""
    if k != 1:
        raise NotImplementedError()
    op = PrecondOperator(A, delta, R, upper_tri)
    return op, op.precond, op.precond_t


FAST_SVD = os.environ.get("PLA_FAST_SVD", "1") != "0"
FAST_SVD_MIN_N = 512          # below this the cuSOLVER SVD is cheap anyway
FAST_SVD_MAX_COND = 1e4       # Gram-based singular vectors carry a relative error ~ eps * cond^2


def _svd_via_gram(A_ske):
    """Thin SVD of a numerically well-conditioned matrix from the symmetric eigendecomposition of its Gram
    matrix: G = A'A = V L V' (cuSOLVER syevd, several times faster than gesvd), B = A V (DMMA GEMM),
    sigma_i = |B e_i| (norms of computed columns, not sqrt(L): no squaring of the small values),
    U = B / sigma.  Returns None unless sigma_min / sigma_max > 1 / FAST_SVD_MAX_COND, in which case the
    vectors are accurate to ~eps * cond^2 <= 2e-8 and no rank decision is involved."""
    G = K.gemm(A_ske, A_ske, transa=True)
    _, V = torch.linalg.eigh(G)
    V = V.flip(1).contiguous()                       # descending order, like an SVD
    B = K.gemm(A_ske, V)
    sigma = torch.linalg.vector_norm(B, dim=0)
    smin, smax = float(sigma.min()), float(sigma.max())
    if not (smin > 0.0 and smin * FAST_SVD_MAX_COND > smax):
        return None
    return V, B / sigma, sigma


def inverse_if_well_conditioned(R_tri, max_cond=None):
    """For an upper-triangular R (the Householder factor of the sketch): the explicit inverse X = R^{-1} if
    cond_2(R) is safely below ``max_cond`` (default FAST_SVD_MAX_COND), else None.

    Why: with R = U_r diag(sigma) Vh the reference's SVD preconditioner is M = V / sigma (preconditioning.py:70-79)
    and R^{-1} = M U_r^T differs from it by an orthogonal factor on the right, so LSQR on A R^{-1} and on A M produce
    the same x = M z and the same |(A M)^T r| history; when the sketch has full numerical rank the n x n
    eigen/singular value decomposition (the only replicated O(n^3) cuSOLVER call of SAP2: 0.09 s at n = 4096) is
    not needed.  cond_2 is estimated by 8 power iterations on R^T R and on X^T X (the decision has eight orders of
    magnitude of slack: rank deficiency means cond > 1 / (n eps) ~ 1e12, the bound asked for is 1e4)."""
    max_cond = FAST_SVD_MAX_COND if max_cond is None else max_cond
    n = R_tri.shape[0]
    X = K.trtri_upper(R_tri)
    if not bool(torch.isfinite(X).all()):
        return None
    g = torch.Generator(device=R_tri.device).manual_seed(1234)
    v0 = torch.randn(n, dtype=F64, device=R_tri.device, generator=g)

    def top_sv(Mx):
        v = v0 / torch.linalg.vector_norm(v0)
        s = v.new_zeros(())
        for _ in range(8):
            w = Mx.T @ (Mx @ v)
            s = torch.linalg.vector_norm(w)
            v = w / s
        return math.sqrt(float(s))

    Rt = torch.triu(R_tri)
    smax, inv_smin = top_sv(Rt), top_sv(X)
    if not (math.isfinite(smax) and math.isfinite(inv_smin)) or smax * inv_smin * 4.0 > max_cond:
        return None
    return X


def svd_right_precond(A_ske, exact=False):
    """parla/comps/preconditioning.py:70-79.  ``exact=True`` forces the true SVD (callers that use U beyond
    the presolve, or that need the rank decision).  The small dense SVD is cuSOLVER glue; for
    n >= 512 and a well-conditioned sketch it is replaced by the Gram/eigh route above (same M, U, sigma, Vh
    up to rotations inside clusters of equal singular values, which no caller can observe)."""
    A_ske = A_ske.contiguous()
    n = A_ske.shape[1]
    if FAST_SVD and not exact and n >= FAST_SVD_MIN_N and A_ske.shape[0] >= n:
        fast = _svd_via_gram(A_ske)
        if fast is not None:
            V, U, sigma = fast
            return (V / sigma).contiguous(), U, sigma, V.T
    # driver 'gesvd' (QR iteration): 1e-14 reconstruction error and ~2.4x faster than the Jacobi default here
    U, sigma, Vh = torch.linalg.svd(A_ske, full_matrices=False, driver='gesvd')
    eps = torch.finfo(F64).eps
    rank = int(torch.count_nonzero(sigma > sigma[0] * A_ske.shape[1] * eps))
    Vh, U, sigma = Vh[:rank, :], U[:, :rank], sigma[:rank]
    M = (Vh.T / sigma).contiguous()
    return M, U, sigma, Vh
