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
        # small dense SVD of the (k+over) x n factor: cuSOLVER glue
        U, s, Vh = torch.linalg.svd(B.contiguous(), full_matrices=False)
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
