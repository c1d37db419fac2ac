"""Randomized interpolative and CUR decompositions.  Mirrors parla/drivers/interpolative.py: ``OneSidedID``
(:10-85), ``OSID1`` (:97-134), ``OSID2`` (:146-176), ``TwoSidedID`` (:179-244), ``TSID1`` (:257-278),
``CURDecomposition`` (:281-345), ``CUR1`` (:358-387) and the procedural ``osid1 / osid2 / tsid1 / cur1``.

The passes over A (sketching, power iteration) are DMMA GEMMs; the column-pivoted QR runs on the
(k + over)-row sketch or on the k selected rows / columns (comps/interpolative.py); the pseudo-inverse
applications are tall full-rank least-squares solves (cuSOLVER glue through ``torch.linalg.lstsq``).
A must be resident on one GPU (no row-sharded form)."""
import numpy as np
import torch

from ..comps import interpolative as id_comps
from ..comps.interpolative import qrcp, sketch_for_axis
from ..parallel import RowSharded
from ..utils import linalg_wrappers as ulaw


def _one_gpu(A):
    if isinstance(A, RowSharded):
        raise NotImplementedError("interpolative decompositions need the matrix on one GPU")


class OneSidedID:
    """``__call__(A, k, over, axis, rng) -> (coefficient matrix, skeleton indices)``: a RowID (axis=0),
    A ~ Z @ A[Is, :] with Z[Is, :] = I, or a ColumnID (axis=1), A ~ A[:, Js] @ X with X[:, Js] = I."""

    def __call__(self, A, k, over, axis, rng):
        raise NotImplementedError()

    exec = __call__


class OSID1(OneSidedID):
    """Sketch A to (k + over) columns / rows, then ID of the sketch (interpolative.py:97-134)."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, over, axis, rng):
        _one_gpu(A)
        rng = np.random.default_rng(rng)
        if axis not in (0, 1):
            raise ValueError()
        Y = sketch_for_axis(self.sk_op, A, k, over, axis, rng)
        return id_comps.qrcp_osid(Y, k, axis=axis)

    exec = __call__


class OSID2(OneSidedID):
    """Skeleton indices from the QRCP of a sketch, coefficients by a pseudo-inverse (interpolative.py:146-176)."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, over, axis, rng):
        _one_gpu(A)
        rng = np.random.default_rng(rng)
        if axis == 0:
            Y = sketch_for_axis(self.sk_op, A, k, over, 0, rng)
            Is = qrcp(Y.T, k)[1][:k]
            X = ulaw.apply_pinv_on_right(A, operator=A[Is, :])
            return X, Is
        elif axis == 1:
            Y = sketch_for_axis(self.sk_op, A, k, over, 1, rng)
            Js = qrcp(Y, k)[1][:k]
            Z = ulaw.apply_pinv_on_left(A, operator=A[:, Js])
            return Z, Js
        else:
            raise ValueError()

    exec = __call__


class TwoSidedID:
    """``__call__(A, k, over, rng) -> (Z, Is, X, Js)`` with A ~ Z @ A[Is, :][:, Js] @ X."""

    def __call__(self, A, k, over, rng):
        raise NotImplementedError()

    exec = __call__


class TSID1(TwoSidedID):
    """One-sided ID on the long axis, then a deterministic ID of the skeleton (interpolative.py:257-278)."""

    def __init__(self, osid: OneSidedID):
        self.osid = osid

    def __call__(self, A, k, over, rng):
        _one_gpu(A)
        rng = np.random.default_rng(rng)
        if A.shape[0] > A.shape[1]:
            X, Js = self.osid(A, k, over, axis=1, rng=rng)
            Z, Is = id_comps.qrcp_osid(A[:, Js], k, axis=0)
        else:
            Z, Is = self.osid(A, k, over, axis=0, rng=rng)
            X, Js = id_comps.qrcp_osid(A[Is, :], k, axis=1)
        return Z, Is, X, Js

    exec = __call__


class CURDecomposition:
    """``__call__(A, k, over, rng) -> (Js, U, Is)`` with A ~ A[:, Js] @ U @ A[Is, :]."""

    def __call__(self, A, k, over, rng):
        raise NotImplementedError()

    exec = __call__


class CUR1(CURDecomposition):
    """interpolative.py:358-387."""

    def __init__(self, osid: OneSidedID):
        self.osid = osid

    def __call__(self, A, k, over, rng):
        _one_gpu(A)
        rng = np.random.default_rng(rng)
        if A.shape[0] > A.shape[1]:
            X, Js = self.osid(A, k, over, axis=1, rng=rng)
            # A \approx A[:, Js] @ X
            Is = qrcp(A[:, Js].T, k)[1][:k]
            U = ulaw.apply_pinv_on_right(X, operator=A[Is, :])
            # U = X (A[Is, :]^\dagger)
            return Js, U, Is
        else:
            Z, Is = self.osid(A, k, over, axis=0, rng=rng)
            # A \approx Z @ A[Is, :]
            Js = qrcp(A[Is, :], k)[1][:k]
            U = ulaw.apply_pinv_on_left(Z, operator=A[:, Js])
            # U = A[:, Js]^\dagger Z
            return Js, U, Is

    exec = __call__


def _rs1(num_pass):
    from ..comps.sketchers import oblivious as osk
    from ..comps.sketchers.aware import RS1
    return RS1(osk.SkOpGA(), num_pass, ulaw.orth, passes_per_stab=1)


def osid1(A, k, over, p, axis, rng):
    """interpolative.py:88-94."""
    return OSID1(_rs1(p - 1))(A, k, over, axis, np.random.default_rng(rng))


def osid2(A, k, over, p, axis, rng):
    """interpolative.py:137-143."""
    return OSID2(_rs1(p - 1))(A, k, over, axis, np.random.default_rng(rng))


def tsid1(A, k, over, p, rng):
    """interpolative.py:247-254."""
    return TSID1(OSID1(_rs1(p - 1)))(A, k, over, np.random.default_rng(rng))


def cur1(A, k, over, p, rng):
    """interpolative.py:348-355."""
    return CUR1(OSID1(_rs1(p - 2)))(A, k, over, np.random.default_rng(rng))
