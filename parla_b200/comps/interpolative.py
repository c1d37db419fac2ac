"""Interpolative-decomposition building blocks.  Mirrors parla/comps/interpolative.py: ``qrcp_osid`` (:11-67),
the ``RowOrColSelection`` interface (:70-119), ``ROCS1`` (:131-157) and ``rocs1`` (:122-128).

The column-pivoted QR acts on a SKETCH ((k + over) x n or (k + over) x m: a few hundred rows), never on A: it is a
Householder QRCP with LAPACK's pivoting rule (largest remaining partial column norm, ``dlaqp2`` downdating with
re-computation when cancellation is detected) written with tensor operations on whatever device the sketch
lives on; the products with A go through the DMMA GEMM (``distla``).
"""
import math

import numpy as np
import torch

from .. import distla

F64 = torch.float64


def qrcp(Y, k=None):
    """Column-pivoted Householder QR of Y (M x N): returns (R, J) with Y[:, J] = Q R, R the min(M, N) x N upper
    trapezoidal factor and J the pivot order -- what ``scipy.linalg.qr(Y, mode='economic', pivoting=True)[1:]``
    returns (LAPACK dgeqp3).  Only the first ``k`` steps are taken when ``k`` is given (J[:k] and R[:k] valid)."""
    W = Y.clone().to(F64)
    M, N = W.shape
    steps = min(M, N) if k is None else min(k, M, N)
    J = torch.arange(N, device=W.device)
    norms = torch.linalg.vector_norm(W, dim=0)
    ref = norms.clone()
    tol3z = math.sqrt(torch.finfo(F64).eps)
    for j in range(steps):
        p = j + int(torch.argmax(norms[j:]))                       # first index of the largest partial norm
        if p != j:
            W[:, [j, p]] = W[:, [p, j]]
            J[[j, p]] = J[[p, j]]
            norms[p], ref[p] = norms[j], ref[j]
        x = W[j:, j]
        alpha = float(x[0])
        xnorm = float(torch.linalg.vector_norm(x[1:])) if M - j > 1 else 0.0
        if xnorm == 0.0:
            tau, beta = 0.0, alpha                                 # H = I (dlarfg)
        else:
            beta = -math.copysign(math.hypot(alpha, xnorm), alpha)
            tau = (beta - alpha) / beta
            v = x / (alpha - beta)
            v[0] = 1.0
            if j + 1 < N:
                C = W[j:, j + 1:]
                C -= torch.outer(tau * v, v @ C)                   # apply H = I - tau v v' to the trailing columns
        W[j, j] = beta
        W[j + 1:, j] = 0.0
        if j + 1 < N:                                              # partial column norm downdate (dlaqp2)
            rest = norms[j + 1:]
            nz = rest != 0
            t = torch.where(nz, W[j, j + 1:].abs() / torch.where(nz, rest, torch.ones_like(rest)), torch.zeros_like(rest))
            t = torch.clamp((1 + t) * (1 - t), min=0.0)
            t2 = t * (rest / torch.where(ref[j + 1:] != 0, ref[j + 1:], torch.ones_like(rest))) ** 2
            redo = nz & (t2 <= tol3z)
            new = rest * torch.sqrt(t)
            if bool(redo.any()):
                fresh = torch.linalg.vector_norm(W[j + 1:, j + 1:], dim=0) if j + 1 < M else torch.zeros_like(rest)
                new = torch.where(redo, fresh, new)
                ref[j + 1:] = torch.where(redo, fresh, ref[j + 1:])
            norms[j + 1:] = torch.where(nz, new, rest)
    r = min(M, N)
    return torch.triu(W[:r, :]), J


def qrcp_osid(Y, k, axis):
    """comps/interpolative.py:11-67.  axis=1: (X, Js) with Y ~ Y[:, Js] @ X and X[:, Js] = I;
    axis=0: (Z, Is) with Y ~ Z @ Y[Is, :]."""
    if axis == 1:
        R, J = qrcp(Y)
        T = torch.linalg.solve_triangular(R[:k, :k], R[:k, k:], upper=True)
        X = torch.zeros(k, Y.shape[1], dtype=F64, device=Y.device)
        X[:, J] = torch.cat((torch.eye(k, dtype=F64, device=Y.device), T), dim=1)
        return X, J[:k]
    elif axis == 0:
        X, Is = qrcp_osid(Y.T, k, axis=1)
        return X.T.contiguous(), Is
    else:
        raise ValueError()


class RowOrColSelection:
    """Interface of comps/interpolative.py:70-119: ``__call__(A, k, over, axis, rng) -> indices``."""

    def __call__(self, A, k, over, axis, rng):
        raise NotImplementedError()

    exec = __call__


def sketch_for_axis(sk_op, A, k, over, axis, rng):
    """The sketch whose QRCP picks the skeleton: ``A @ S`` (axis 0, m x (k+over)) or ``S' @ A`` with
    ``S = sk_op(A', k + over)`` (axis 1, (k+over) x n)  -- interpolative.py:143-153."""
    if axis == 0:
        return distla.mm(A, sk_op(A, k + over, rng))
    if axis == 1:
        S = sk_op(A.T, k + over, rng)                              # m x (k + over)
        return distla.mm_t(S, A)
    raise ValueError()


class ROCS1(RowOrColSelection):
    """Sketch + QRCP skeleton (comps/interpolative.py:131-157)."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, over, axis, rng):
        rng = np.random.default_rng(rng)
        Y = sketch_for_axis(self.sk_op, A, k, over, axis, rng)
        return qrcp(Y.T if axis == 0 else Y, k)[1][:k]

    exec = __call__


def rocs1(A, k, over, p, axis, rng):
    """comps/interpolative.py:122-128."""
    from .sketchers import oblivious as osk
    from .sketchers import aware as ask
    from ..utils import linalg_wrappers as ulaw
    rng = np.random.default_rng(rng)
    return ROCS1(ask.RS1(osk.SkOpGA(), p - 1, ulaw.orth, passes_per_stab=1))(A, k, over, axis, rng)
