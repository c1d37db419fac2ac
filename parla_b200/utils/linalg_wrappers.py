"""parla/utils/linalg_wrappers.py on the device.

``orth`` (:6-7) is the Q factor of an economic Householder QR (hand-written kernels; a Householder TSQR when
the argument is ``RowSharded``).  The LU / pseudo-inverse helpers (:10-35) act on small dense matrices and are
cuSOLVER glue through ``torch.linalg``."""
import torch

from .. import distla
from ..parallel import RowSharded


def orth(S):
    return distla.orth(S)


def lu_stabilize(S):
    """linalg_wrappers.py:10-12: the row-permuted unit lower-trapezoidal factor ``P @ L`` of ``S = P L U``."""
    if isinstance(S, RowSharded):
        raise NotImplementedError("lu_stabilize needs the matrix on one GPU")
    P, L, _ = torch.linalg.lu(S)
    return (P @ L).contiguous()


def lupt(M):
    """linalg_wrappers.py:15-18: factor M = L @ U @ P.T (equivalently M @ P = L @ U)."""
    P, L, U = torch.linalg.lu(M.T.contiguous())
    return U.T.contiguous(), L.T.contiguous(), P


def lup(M):
    """linalg_wrappers.py:21-24: factor M = L @ U @ P."""
    P, L, U = torch.linalg.lu(M.T.contiguous())
    return U.T.contiguous(), L.T.contiguous(), P.T.contiguous()


def apply_pinv_on_left(target, operator):
    """linalg_wrappers.py:27-30: pinv(operator) @ target  (tall full-rank ``operator``: QR-based least squares)."""
    return torch.linalg.lstsq(operator, target).solution


def apply_pinv_on_right(target, operator):
    """linalg_wrappers.py:33-36: target @ pinv(operator)."""
    return torch.linalg.lstsq(operator.T.contiguous(), target.T.contiguous()).solution.T.contiguous()
