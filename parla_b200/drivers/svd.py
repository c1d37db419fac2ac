"""Randomized SVD driver.  Mirrors parla/drivers/svd.py: interface (:10-55), ``SVD1`` (:126-176)."""
import numpy as np
import torch

from .. import distla
from ..comps.qb import QBDecomposer


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
