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


def _pinv_apply(operator, target):
    """pinv(operator) @ target through a thin SVD with LAPACK gelsd's cut-off (singular values below
    eps * sigma_max are treated as zero): the minimum-norm least-squares solution also when ``operator`` is
    rank deficient, which torch.linalg.lstsq's only CUDA driver ('gels', full rank assumed) does not give."""
    U, s, Vh = torch.linalg.svd(operator, full_matrices=False)
    cut = torch.finfo(operator.dtype).eps * (s[0] if s.numel() else 0.0)
    sinv = torch.where(s > cut, 1.0 / s, torch.zeros_like(s))
    return Vh.T @ (sinv.unsqueeze(1) * (U.T @ target))


def apply_pinv_on_left(target, operator):
    """linalg_wrappers.py:27-30: pinv(operator) @ target  (scipy ``lstsq`` / gelsd semantics)."""
    return _pinv_apply(operator, target).contiguous()


def apply_pinv_on_right(target, operator):
    """linalg_wrappers.py:33-36: target @ pinv(operator)."""
    return _pinv_apply(operator.T.contiguous(), target.T.contiguous()).T.contiguous()
