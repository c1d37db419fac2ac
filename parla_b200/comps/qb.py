"""QB decompositions.  Mirrors parla/comps/qb.py: interface (:241-283), ``QB1`` (:286-352),
``QB2`` (:355-482) and ``project_out`` (:603-613).

QB2 preallocates Q and B (the reference re-stacks them every block, :473-474) and deflates A in
place with one rank-blk DMMA update (:475).
"""
import warnings

import numpy as np
import torch

from .. import distla
from .. import kernels as K
from ..parallel import RowSharded
from .rangefinders import RangeFinder

F64 = torch.float64


class QBDecomposer:

    def __call__(self, A, k, tol, rng):
        raise NotImplementedError()

    exec = __call__


class QB1(QBDecomposer):

    TOL_CONTROL = 'unknown'

    def __init__(self, rf: RangeFinder):
        self.rangefinder = rf

    def __call__(self, A, k, tol, rng):
        assert k > 0                                               # qb.py:344-348
        assert k <= min(A.shape)
        if not np.isnan(tol):
            assert tol >= 0
            assert tol < 1
        rng = np.random.default_rng(rng)
        Q = self.rangefinder(A, k, tol, rng)
        B = distla.mm_t(Q, A)                                      # :351
        return Q, B

    exec = __call__


def project_out(Qi, Q, as_list=False):
    """Qi - Q (Q^T Qi)   (qb.py:603-613)."""
    if as_list:
        raise NotImplementedError()
    if Q.shape[1] == 0:
        return Qi
    C = distla.mm_t(Q, Qi)
    return distla.sub_outer_(distla.clone(Qi), Q, C)


class QB2(QBDecomposer):

    TOL_CONTROL = 'full'

    def __init__(self, rf: RangeFinder, blk: int, overwrite_a: bool):
        self.rangefinder = rf
        self.blk = blk
        self.overwrite_a = overwrite_a

    def __call__(self, A, k, tol, rng):
        if not self.overwrite_a:                                   # qb.py:442-443
            A = distla.clone(A)
        assert k > 0
        small_dim = min(A.shape)
        if not k <= small_dim:                                     # :445-452
            msg = f"""
            The target rank k = {k} is larger than min({tuple(A.shape)}).
            We will proceed with target rank k = {small_dim}.
            """
            k = small_dim
            warnings.warn(msg)
        assert k <= min(A.shape)
        use_tol = not np.isnan(tol) and tol > 0                    # :454
        if use_tol:
            sq_norm_A = float(distla.sumsq_all(A))
            abs_sq_tol = sq_norm_A * tol ** 2
        rng = np.random.default_rng(rng)
        m, n = A.shape
        sharded = isinstance(A, RowSharded)
        m_loc = A.local.shape[0] if sharded else m
        Qbuf = torch.empty(m_loc, k, dtype=F64, device=A.device)
        B = torch.empty(k, n, dtype=F64, device=A.device)

        def q_view(c0, c1):
            v = Qbuf[:, c0:c1]
            return RowSharded(v, A.row_offset, A.m_global, A.group) if sharded else v

        cols = 0
        blk = self.blk
        while True:                                                # :463-481
            if cols + blk > k:
                blk = k - cols  # final block
            Qi = self.rangefinder(A, blk, np.nan, rng)
            Qi = project_out(Qi, q_view(0, cols))
            Qi = distla.orth(Qi)
            Bi = distla.mm_t(Qi, A, out=B[cols:cols + blk])
            Qbuf[:, cols:cols + blk] = distla.local(Qi)
            cols += blk
            distla.sub_outer_(A, Qi, Bi)                           # A -= Qi @ Bi
            if use_tol:
                sq_norm_A = sq_norm_A - float(K.sumsq(Bi.reshape(-1)))
                if sq_norm_A <= abs_sq_tol:
                    break
            if cols >= k:
                break
        return q_view(0, cols), B[:cols]

    exec = __call__
