"""Dense products and orthogonalisation that work on plain device tensors and on ``RowSharded``
tall operands alike (low-rank path).

Row-sharded rule set: ``A @ S`` is local (no communication), ``A^T @ Y`` and ``Q^T @ A`` are local
DMMA GEMMs followed by an all-reduce(sum) of the small n x k / k x n block, the rank-k deflation
``A -= Q B`` is local, and ``orth`` of a row-sharded tall matrix is a Householder TSQR (local QR,
all-gather of the k x k R factors, QR of the stacked R's, local GEMM).
"""
import os

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


CHOLQR_MIN_ROWS = 1 << 15       # below this the Householder QR is cheap (and its Q matches LAPACK's signs)
CHOLQR_MAX_DEFECT = 0.05        # |Q1'Q1 - I|_max after the first round (~ eps cond(Y)^2): second round then reaches eps


def _cholqr2(Y, group):
    """Orthonormal basis of range(Y) for a TALL, numerically well-conditioned Y (m >> k) by two rounds of
    Cholesky QR: Gram matrix (DMMA GEMM, all-reduced over row shards) -> k x k Cholesky (cuSOLVER glue) ->
    Q = Y R^{-1} (explicit triangular inverse + DMMA GEMM).  Four GEMM passes over Y instead of the Householder
    panel sweeps (2^20 x 512: ~70 ms instead of 232 ms).  The first round loses orthogonality as eps cond(Y)^2;
    that defect is MEASURED (|Q1'Q1 - I|) and the routine returns None -- the caller falls back to Householder --
    unless it is small enough for the second round to restore orthogonality to machine precision (cond(Y) up to
    ~1e7).  Rank-deficient or ill-conditioned Y (Cholesky breakdown) also returns None.  The stabiliser of the
    power iteration (aware.py:160-184) keeps its operand well conditioned, which is the case this is for."""
    k = Y.shape[1]
    eye = torch.eye(k, dtype=F64, device=Y.device)
    Q = Y
    for rnd in range(2):
        G = allreduce_(K.gemm(Q, Q, transa=True), group)
        if rnd == 1 and not float((G - eye).abs().max()) < CHOLQR_MAX_DEFECT:
            return None
        R, info = torch.linalg.cholesky_ex(G, upper=True)
        if int(info) != 0 or not bool(torch.isfinite(R).all()):
            return None
        Q = K.gemm(Q, K.trtri_upper(R.contiguous()))
    return Q


def orth(Y):
    """Orthonormal basis of range(Y) (``la.qr(Y, mode='economic')[0]``, utils/linalg_wrappers.py:6-7, up to the signs
    of the columns): CholeskyQR2 for tall well-conditioned operands, Householder QR otherwise (a Householder TSQR
    when row-sharded)."""
    Yl = Y.local if isinstance(Y, RowSharded) else Y
    group = _group(Y) if isinstance(Y, RowSharded) else None
    rows = Y.shape[0]
    if (os.environ.get("PLA_CHOLQR", "1") != "0" and Yl.dim() == 2 and rows >= CHOLQR_MIN_ROWS
            and rows >= 16 * Yl.shape[1] and Yl.shape[0] >= Yl.shape[1]):
        Q = _cholqr2(Yl if Yl.stride(1) == 1 else Yl.contiguous(), group)
        if os.environ.get("PLA_TRACE_ORTH"):
            import sys
            print(f"[orth] {tuple(Yl.shape)}: {'CholeskyQR2' if Q is not None else 'Householder (fallback)'}", file=sys.stderr)
        if Q is not None:
            return _like(Q, Y) if isinstance(Y, RowSharded) else Q
    if not isinstance(Y, RowSharded):
        return K.qr_economic(Y)[0]
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


# ------------------------------------------------------------------------------------------------------------------
# QR of the REPLICATED d x (n + 1) sketch [A_ske | b_ske] with the trailing columns spread over the ranks.
QR_DIST_MIN_BLOCKS_PER_RANK = 2
QR_DIST_MAX_ROWS = 148 * 160        # capacity of the cooperative block kernel (csrc/qr.cu: QB_RPC_MAX rows per SM)


def geqrf_distributed_ok(d, n, group):
    """Worth it (and possible) when every rank gets at least two 128-column blocks and a panel fits the
    cooperative kernel; otherwise callers use the replicated K.geqrf."""
    if group is None or not dist.is_initialized():
        return False
    world = dist.get_world_size(group)
    return world > 1 and d <= QR_DIST_MAX_ROWS and n >= QR_DIST_MIN_BLOCKS_PER_RANK * K.QR_BLOCK * world


class _DeviceBlockQR:
    """The three block-level kernel calls geqrf_distributed is built from (tests substitute a LAPACK stand-in to
    check the ownership / packing / collective logic on CPU with gloo)."""
    workspace = staticmethod(K.qr_block_workspace)
    factor = staticmethod(K.qr_factor_block)
    apply = staticmethod(K.qr_apply_block)


def geqrf_distributed(W, n, group, _backend=_DeviceBlockQR):
    """In-place Householder QR of the first n columns of W (d x (n + 1), identical on every rank; the last column is
    carried along as a right-hand side), replacing ``la.qr(A_ske)`` + ``Q.T @ b_ske`` (least_squares.py:311,315)
    when A is row-sharded.  After the sketch all-reduce every rank would otherwise repeat the whole factorisation --
    the Amdahl term of strong scaling (44 % of an 8-GPU solve at 16384 x 4097 in round 1).  Here:

      * 128-column blocks are dealt to the ranks round-robin; a rank keeps its blocks PACKED (d x n/P, contiguous)
        plus its own copy of the right-hand-side column;
      * per block: the owner broadcasts the current panel (d x 128, NCCL over NVLink), EVERY rank factors it with the
        cooperative block kernel (identical bits everywhere: same input, deterministic kernel) and applies the block
        reflector to its own packed trailing columns only -- the tensor-core trailing updates, ~70 % of the
        factorisation, are divided by the number of ranks;
      * the R factor (rows 0..n-1 of the packed blocks) is all-gathered at the end.

    On return ``triu(W[:n, :n]) = R`` and ``W[:n, n] = (Q^T [b_ske; 0])[:n]`` on every rank, as after K.geqrf (the
    reflectors below the diagonal are NOT gathered: the solvers never form Q)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    d = W.shape[0]
    nb = K.QR_BLOCK
    nblk = (n + nb - 1) // nb
    mine = list(range(rank, nblk, world))                      # global indices of my blocks
    widths = [min(nb, n - nb * j) for j in mine]
    ncl = sum(widths)
    C = torch.empty(d, ncl + 1, dtype=F64, device=W.device)    # packed owned columns | rhs column
    off = 0
    for j, w in zip(mine, widths):
        C[:, off:off + w].copy_(W[:, nb * j:nb * j + w])
        off += w
    C[:, ncl].copy_(W[:, n])
    n_layout = max(ncl + 1, nb)
    ws = _backend.workspace(W.device, d, n_layout)
    panel = torch.empty(d, nb, dtype=F64, device=W.device)
    tau = torch.empty(nblk * nb, dtype=F64, device=W.device)
    ranks = dist.get_process_group_ranks(group)
    for k in range(nblk):
        owner, kl = k % world, k // world
        wk = min(nb, n - nb * k)
        if rank == owner:
            panel[:, :wk].copy_(C[:, nb * kl:nb * kl + wk])
        dist.broadcast(panel, src=ranks[owner], group=group)
        _backend.factor(panel, nb * k, 0, wk, tau[nb * k:], k, n_layout, ws)
        if rank == owner:
            C[:, nb * kl:nb * kl + wk].copy_(panel[:, :wk])
        first = 0 if k < rank else (k - rank) // world + 1     # my blocks with global index <= k
        c0 = nb * first if first < len(mine) else ncl
        _backend.apply(d, nb * k, wk, tau[nb * k:], C[:, c0:], n_layout, ws)
    # assemble R (and keep the rhs column, which every rank carried through all the reflectors)
    per = (nblk + world - 1) // world * nb
    send = torch.zeros(n, per, dtype=F64, device=W.device)
    send[:, :ncl].copy_(C[:n, :ncl])
    recv = torch.empty(world * n, per, dtype=F64, device=W.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, n, per)
    for r in range(world):
        off = 0
        for j in range(r, nblk, world):
            w = min(nb, n - nb * j)
            W[:n, nb * j:nb * j + w].copy_(recv[r, :, off:off + w])
            off += w
    W[:n, n].copy_(C[:n, ncl])
    return tau[:n]
