"""Randomized eigendecomposition driver.  Mirrors parla/drivers/evd.py: interface (:171-208),
``EVD1`` (:211-289; A symmetric)."""
import numpy as np
import torch
import torch.distributed

from .. import distla
from .. import kernels as K
from ..parallel import RowSharded, allreduce_
from ..comps.qb import QBDecomposer


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
