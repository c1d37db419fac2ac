"""Randomized eigendecomposition drivers.  Mirrors parla/drivers/evd.py: interface (:171-208),
``EVD1`` (:211-289; A symmetric) and ``EVD2`` (:290-381; A symmetric PSD, Nystrom)."""
import math
import warnings

import numpy as np
import torch
import torch.distributed

from .. import distla
from .. import kernels as K
from ..parallel import RowSharded, allreduce_
from ..comps.qb import QBDecomposer


def evd1(A, k, tol, over, inner_num_pass, block_size, rng):
    """evd.py:16-96 (note: tol is halved here and again inside EVD1, as in the reference)."""
    from ..comps.sketchers import oblivious
    from ..comps.sketchers.aware import RS1
    from ..comps.rangefinders import RF1
    from ..comps.qb import QB2
    from ..utils import linalg_wrappers as ulaw
    assert inner_num_pass >= 2
    rng = np.random.default_rng(rng)
    rso_ = RS1(oblivious.SkOpGA(), inner_num_pass - 2, ulaw.orth, 1)
    return EVD1(QB2(RF1(rso_), block_size, overwrite_a=False))(A, k, tol / 2, over, rng)


def evd2(A, k, over, num_passes, rng):
    """evd.py:99-164."""
    from ..comps.sketchers import oblivious
    from ..comps.sketchers.aware import RS1
    from ..utils import linalg_wrappers as ulaw
    assert num_passes >= 1
    rng = np.random.default_rng(rng)
    return EVD2(RS1(oblivious.SkOpGA(), num_passes - 1, ulaw.orth, 1))(A, k, np.nan, over, rng)


class EVDecomposer:

    def __call__(self, A, k, tol, over, rng):
        raise NotImplementedError()

    exec = __call__


class EVD1(EVDecomposer):

    TOL_CONTROL = 'unknown'

    def __init__(self, qb: QBDecomposer):
        self.qb = qb

    def __call__(self, A, k, tol, over, rng):
        assert k > 0                                               # evd.py:270-274
        assert k <= min(A.shape)
        if not np.isnan(tol):
            assert tol >= 0
            assert tol < np.inf
        rng = np.random.default_rng(rng)
        Q, B = self.qb(A, k + over, tol / 2, rng)                  # :276
        if isinstance(Q, RowSharded):                              # B replicated (k x n), Q row-sharded (n x k)
            lo = Q.row_offset
            C = K.gemm(B[:, lo:lo + Q.local.shape[0]], Q.local)
            allreduce_(C, Q.group if Q.group is not None else torch.distributed.group.WORLD)
        else:
            C = K.gemm(B, Q)                                       # :278
        lamb, U = torch.linalg.eigh(C)                             # :279 small dense: cuSOLVER glue
        alamb = torch.abs(lamb)
        d = Q.shape[1]
        r = min(k, d, int(torch.count_nonzero(alamb > 10 * np.finfo(float).eps)))
        I = torch.argsort(-alamb, stable=True)[:r]                 # :284
        U = U[:, I]
        lamb = lamb[I]
        V = distla.mm(Q, U.contiguous())                           # :288
        return V, lamb

    exec = __call__


class EVD2(EVDecomposer):
    """Rank-k truncation of the regularised Nystrom approximation (A S)(S'A S)^+ (A S)' of a symmetric PSD
    matrix (evd.py:290-381; Tropp, Yurtsever, Udell, Cevher 2017, Algorithm 3)."""

    TOL_CONTROL = 'none'

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, tol, over, rng):
        assert k > 0                                               # evd.py:352-354
        n = A.shape[0]
        assert k < n
        if not np.isnan(tol):
            msg = """
            This EVDecomposer implementation cannot directly control
            approximation error. Parameter "tol" is being ignored.
            """
            warnings.warn(msg)
        if isinstance(A, RowSharded):
            raise NotImplementedError("EVD2 is implemented for a matrix resident on one GPU")
        rng = np.random.default_rng(rng)
        S = self.sk_op(A, k + over, rng).contiguous()              # n x (k + over)
        Y = K.gemm(A, S)                                           # :363
        nu = math.sqrt(n) * np.finfo(float).eps * math.sqrt(float(K.sumsq(Y.reshape(-1))))   # :365
        Y.add_(S, alpha=nu)                                        # temporary regularisation (:367)
        R = torch.linalg.cholesky(K.gemm(S, Y, transa=True), upper=True)      # small dense: cuSOLVER glue
        Bm = K.gemm(Y, K.trtri_upper(R))                           # Y R^-1 = (R^-T Y')'   (:370)
        Qb, Rb = K.qr_economic(Bm)                                 # thin SVD of the tall B through its QR
        U, sigma, _ = torch.linalg.svd(Rb, full_matrices=False)
        sig2 = (sigma * sigma).cpu().numpy()
        r = min([k] + [i for i in range(k - 1) if sig2[i + 1] <= nu])          # :373-377
        V = K.gemm(Qb, U[:, :r].contiguous())
        lamb = (sigma * sigma)[:r] - nu
        return V, lamb

    exec = __call__
