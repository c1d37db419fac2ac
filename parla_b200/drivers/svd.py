"""Randomized SVD driver.  Mirrors parla/drivers/svd.py: interface (:10-55), ``SVD1`` (:126-176)."""
import numpy as np
import torch

from .. import distla
from ..comps.qb import QBDecomposer


def svd1(A, k, over, tol, inner_num_pass, block_size, rng):
    """svd.py:58-123: RS1 (Gaussian, inner_num_pass - 2 power-iteration passes, QR-stabilised) -> RF1 ->
    QB2(block_size) -> SVD1."""
    from ..comps.sketchers import oblivious
    from ..comps.sketchers.aware import RS1
    from ..comps.rangefinders import RF1
    from ..comps.qb import QB2
    from ..utils import linalg_wrappers as ulaw
    rng = np.random.default_rng(rng)
    rso_ = RS1(oblivious.SkOpGA(), inner_num_pass - 2, ulaw.orth, 1)
    return SVD1(QB2(RF1(rso_), block_size, overwrite_a=False))(A, k, tol, over, rng)


def _svd_of_wide(B):
    """Thin SVD of the (k + over) x n factor B (``la.svd(B, full_matrices=False)``, svd.py:164).  For a wide B
    (n >= 4 rows) it goes through the Householder QR of B^T (hand-written kernels): B^T = Q_b R_b, R_b^T = U s W^T
    (the only cuSOLVER call left is this k x k SVD) and Vh = (Q_b W)^T -- backward stable like the direct SVD, and
    the 512 x 16384 factor of BASELINE configs[3] no longer goes through cuSOLVER's gesvd as a whole."""
    from .. import kernels as K
    kk, n = B.shape
    if n < 4 * kk or kk < 32:
        return torch.linalg.svd(B, full_matrices=False)
    Qb, Rb = K.qr_economic(B.T.contiguous())                   # n x kk, kk x kk
    U, s, Wh = torch.linalg.svd(Rb.T.contiguous(), full_matrices=False)
    Vh = K.gemm(Wh.contiguous(), Qb, transb=True)              # (Q_b W)^T = W^T Q_b^T
    return U, s, Vh


class SVDecomposer:

    def __call__(self, A, k, tol, over, rng):
        raise NotImplementedError()

    exec = __call__


class SVD1(SVDecomposer):

    def __init__(self, qb: QBDecomposer):
        self.qb = qb

    def __call__(self, A, k, tol, over, rng):
        rng = np.random.default_rng(rng)
        Q, B = self.qb(A, k + over, tol, rng)
        U, s, Vh = _svd_of_wide(B.contiguous())
        if over > 0:                                               # svd.py:165-169
            cutoff = min(k, s.numel())
            U, s, Vh = U[:, :cutoff], s[:cutoff], Vh[:cutoff, :]
        drop = s < 10 * np.finfo(float).eps                        # :170-174
        if bool(drop.any()):
            keep = ~drop
            U, s, Vh = U[:, keep], s[keep], Vh[keep, :]
        U = distla.mm(Q, U.contiguous())                           # :175
        return U, s, Vh

    exec = __call__
