"""Dense products and orthogonalisation that work on plain device tensors and on ``RowSharded``
tall operands alike (low-rank path).

Row-sharded rule set: ``A @ S`` is local (no communication), ``A^T @ Y`` and ``Q^T @ A`` are local
DMMA GEMMs followed by an all-reduce(sum) of the small n x k / k x n block, the rank-k deflation
``A -= Q B`` is local, and ``orth`` of a row-sharded tall matrix is a Householder TSQR (local QR,
all-gather of the k x k R factors, QR of the stacked R's, local GEMM).
"""
import torch
import torch.distributed as dist

from . import kernels as K
from .parallel import RowSharded, allreduce_

F64 = torch.float64


def _like(local, ref):
    return RowSharded(local, ref.row_offset, ref.m_global, ref.group)


def _group(X):
    return X.group if X.group is not None else dist.group.WORLD


def _is_transposed_view(A):
    """True for ``B.T`` of a row-major B (column-major strides): its products are issued on B with the
    transpose flag instead of materialising the transpose."""
    return (isinstance(A, torch.Tensor) and A.dim() == 2 and A.shape[0] > 1 and A.shape[1] > 1
            and A.stride(0) == 1 and A.stride(1) >= A.shape[0])


def mm(A, S, alpha=1.0):
    """A @ S with S replicated.  Row-sharded A gives a row-sharded result."""
    if isinstance(A, RowSharded):
        return _like(K.gemm(A.local, S, alpha=alpha), A)
    if _is_transposed_view(A):
        return K.gemm(A.T, S, transa=True, alpha=alpha)
    return K.gemm(A, S, alpha=alpha)


def mm_t(A, Y, out=None):
    """A^T @ Y.  If both are row-sharded the partial products are summed over the ranks."""
    if isinstance(A, RowSharded) or isinstance(Y, RowSharded):
        if not (isinstance(A, RowSharded) and isinstance(Y, RowSharded)):
            raise ValueError("A^T @ Y needs both operands sharded over the same rows")
        Z = K.gemm(A.local, Y.local, transa=True, out=out)
        return allreduce_(Z, _group(A))
    if _is_transposed_view(A):
        return K.gemm(A.T, Y, out=out)                 # (B')' Y = B Y
    return K.gemm(A, Y, transa=True, out=out)


def sub_outer_(A, Q, B):
    """A -= Q @ B in place (Q row-sharded like A, B replicated)."""
    if isinstance(A, RowSharded):
        K.gemm(Q.local, B, alpha=-1.0, beta=1.0, out=A.local)
    else:
        K.gemm(Q, B, alpha=-1.0, beta=1.0, out=A)
    return A


def sumsq_all(X):
    """Squared Frobenius norm (device scalar), summed over ranks when sharded."""
    if isinstance(X, RowSharded):
        return allreduce_(K.sumsq(X.local.reshape(-1)), _group(X))
    return K.sumsq(X.reshape(-1))


def orth(Y):
    """Orthonormal basis of range(Y): Q factor of a Householder QR (TSQR when row-sharded)."""
    if not isinstance(Y, RowSharded):
        return K.qr_economic(Y)[0]
    group = _group(Y)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    k = Y.local.shape[1]
    if Y.local.shape[0] < k:
        raise ValueError("TSQR needs at least as many local rows as columns on every rank")
    Q1, R1 = K.qr_economic(Y.local)
    stack = torch.empty(world * k, k, dtype=F64, device=R1.device)
    dist.all_gather_into_tensor(stack, R1.contiguous(), group=group)
    Q2, _ = K.qr_economic(stack)                      # identical on every rank
    return _like(K.gemm(Q1, Q2[rank * k:(rank + 1) * k].contiguous()), Y)


def qr(Y):
    """(Q, R) of a tall matrix (Householder; TSQR when row-sharded: R is replicated)."""
    if not isinstance(Y, RowSharded):
        return K.qr_economic(Y)
    group = _group(Y)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    k = Y.local.shape[1]
    if Y.local.shape[0] < k:
        raise ValueError("TSQR needs at least as many local rows as columns on every rank")
    Q1, R1 = K.qr_economic(Y.local)
    stack = torch.empty(world * k, k, dtype=F64, device=R1.device)
    dist.all_gather_into_tensor(stack, R1.contiguous(), group=group)
    Q2, R2 = K.qr_economic(stack)
    return _like(K.gemm(Q1, Q2[rank * k:(rank + 1) * k].contiguous()), Y), R2


def local(X):
    return X.local if isinstance(X, RowSharded) else X


def clone(X):
    return _like(X.local.clone(), X) if isinstance(X, RowSharded) else X.clone()
