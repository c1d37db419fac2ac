"""QB decompositions.  Mirrors parla/comps/qb.py: interface (:241-283), ``QB1`` (:286-352),
``QB2`` (:355-482), ``QB3`` (:484-598) and ``project_out`` (:603-613).

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


def _default_rs1(num_pass):
    from .sketchers import oblivious
    from .sketchers.aware import RS1
    from ..utils import linalg_wrappers as ulaw
    return RS1(oblivious.SkOpGA(), num_pass, ulaw.orth, 1)


def qb(num_passes, A, k, rng):
    """qb.py:16-82: RS1 (num_passes - 2 power-iteration passes) -> RF1 -> QB1."""
    from .rangefinders import RF1
    rng = np.random.default_rng(rng)
    return QB1(RF1(_default_rs1(num_passes - 2)))(A, k, np.nan, rng)


def qb_b(inner_num_pass, blk, overwrite_A, A, k, tol, rng):
    """qb.py:85-167: blocked QB ([YGL:2018, Algorithm 2] up to the differences listed there)."""
    from .rangefinders import RF1
    rng = np.random.default_rng(rng)
    return QB2(RF1(_default_rs1(inner_num_pass - 2)), blk, overwrite_A)(A, k, tol, rng)


def qb_b_pe(num_passes, blk, A, k, tol, rng):
    """qb.py:170-234: pass-efficient blocked QB (QB3)."""
    rng = np.random.default_rng(rng)
    return QB3(_default_rs1(num_passes - 1), blk)(A, k, tol, rng)


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


class QB3(QBDecomposer):
    """Blocked QB from the two sketches G = A S and H = A' G (qb.py:484-598): after those two passes over A
    everything is m x blk / n x blk work (DMMA GEMMs, Householder QR of the blocks, blk x blk solves)."""

    TOL_CONTROL = 'early stopping'

    def __init__(self, sk_op, blk: int):
        self.sk_op = sk_op
        self.blk = blk

    def __call__(self, A, k, tol, rng):
        assert k > 0                                               # qb.py:555-556
        assert k < min(A.shape)
        use_tol = not np.isnan(tol) and tol > 0
        if use_tol:
            sq_norm_A = float(distla.sumsq_all(A))
            abs_sq_tol = sq_norm_A * tol ** 2
        rng = np.random.default_rng(rng)
        blk = self.blk
        S = self.sk_op(A, k, rng)
        if not isinstance(S, torch.Tensor):                        # :568-574
            msg = """
            This implementation requires the sketching routine to return a
            dense matrix, as represented by a device tensor. We received a
            matrix of type %s
            """ % str(type(S))
            raise RuntimeError(msg)
        S = S.contiguous()
        G = distla.mm(A, S)                                        # :575  m x k (row-sharded like A)
        H = distla.mm_t(A, G)                                      # :576  n x k (replicated)
        m, n = A.shape
        sharded = isinstance(A, RowSharded)
        m_loc = A.local.shape[0] if sharded else m
        Qbuf = torch.empty(m_loc, k, dtype=F64, device=A.device)
        B = torch.empty(k, n, dtype=F64, device=A.device)
        Gl = distla.local(G)

        def wrap(v):
            return RowSharded(v, A.row_offset, A.m_global, A.group) if sharded else v

        cols = 0
        for lo in range(0, k, blk):                                # :577-597
            hi = min(lo + blk, k)
            Si = S[:, lo:hi].contiguous()
            Q, Bc = wrap(Qbuf[:, :cols]), B[:cols]
            Yi_loc = Gl[:, lo:hi].contiguous()
            if cols > 0:
                BSi = K.gemm(Bc, Si)                               # cols x b
                K.gemm(Qbuf[:, :cols], BSi, alpha=-1.0, beta=1.0, out=Yi_loc)
            Yi = wrap(Yi_loc)
            Qi, Ri = distla.qr(Yi)
            Qi = project_out(Qi, Q)                                # Qi - Q (Q' Qi)
            Qi, Rihat = distla.qr(Qi)
            Ri = K.gemm(Rihat, Ri)
            Bi = H[:, lo:hi].T.contiguous()
            if cols > 0:
                YtQ = distla.mm_t(Yi, Q)                           # b x cols
                K.gemm(YtQ, Bc, alpha=-1.0, beta=1.0, out=Bi)
                K.gemm(BSi, Bc, transa=True, alpha=-1.0, beta=1.0, out=Bi)
            # Ri' X = Bi  (b x b triangular system with n right-hand sides: small dense glue)
            Bi = torch.linalg.solve_triangular(Ri.T, Bi, upper=False)
            Qbuf[:, cols:cols + (hi - lo)] = distla.local(Qi)
            B[cols:cols + (hi - lo)] = Bi
            cols += hi - lo
            if use_tol:
                sq_norm_A = sq_norm_A - float(K.sumsq(Bi.contiguous().reshape(-1)))
                if sq_norm_A <= abs_sq_tol:
                    break  # early stopping
        return wrap(Qbuf[:, :cols]), B[:cols]

    exec = __call__
