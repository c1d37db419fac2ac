"""Sketching operators for the B200 path.

Mirrors parla/utils/sketching.py:20-31 (gaussian_operator) and :34-80 (sjlt_operator): same call
signatures ``(n_rows, n_cols, rng[, ...])``, but what comes back is a *device operator object*
supporting ``S @ A`` / ``S @ b`` / ``.T`` / ``.shape`` -- a wide Gaussian operator is virtual (its
entries are generated inside the DMMA kernel from a Philox key and never touch HBM), an SJLT is
held in index form plus a destination-major plan.

numpy's PCG64/ziggurat stream cannot be reproduced on a GPU, so the native operators draw ONE
63-bit key from the caller's numpy Generator (advancing it exactly once per call, so nested use
stays deterministic) and derive everything else from Philox4x32-10.  To replay the reference's own
operator (parity tests), wrap it with :func:`as_device_operator`.
"""
import math
import warnings

import numpy as np
import torch

from .. import kernels as K

F64 = torch.float64


def _device(device):
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


_SHARD = None          # (first column, number of columns) of the operator this rank applies


class shard_context:
    """While active, wide operators built by the generators cover only this rank's column range
    (= its rows of A).  Counter-based generation makes the slices consistent across ranks."""

    def __init__(self, col_offset, n_local):
        self.val = (int(col_offset), int(n_local))

    def __enter__(self):
        global _SHARD
        self.prev, _SHARD = _SHARD, self.val

    def __exit__(self, *exc):
        global _SHARD
        _SHARD = self.prev


def _draw_key(rng):
    rng = np.random.default_rng(rng)
    return int(rng.integers(0, 2 ** 63 - 1)), rng


class SketchOperator:
    """Base class of device sketching operators (duck-types what SPO / RS1 need from ``S``)."""
    shape = (0, 0)

    def __matmul__(self, other):
        if isinstance(other, torch.Tensor):
            if other.dim() == 1:
                return self.apply(other.reshape(-1, 1)).reshape(-1)
            return self.apply(other)
        return NotImplemented

    def apply(self, A):                      # S @ A, A (m, n) -> (d, n)
        raise NotImplementedError()

    def sketch_into(self, A, b, out, row_offset=0):
        """out[:d, :n] = S[:, off:off+m] @ A and (b given) out[:d, n] = S[:, off:off+m] @ b."""
        raise NotImplementedError()

    def to_dense(self):
        raise NotImplementedError()

    def column_slice(self, first, count):
        """Operator restricted to columns [first, first+count) (a row shard of A)."""
        raise NotImplementedError()

    def rmatvec(self, v, m_local=None, row_offset=0):
        """S[:, off:off+m_local]^T @ v for ONE vector v (d entries) -> m_local entries
        (``S.T @ v[:d]``, saddlesys.py:291)."""
        raise NotImplementedError()

    @property
    def T(self):
        return self.to_dense().T


class GaussianOperator(SketchOperator):
    """Virtual Gaussian operator G(seed)[:n_rows, :n_cols] * scale  (oracle/philox_ref.py)."""

    def __init__(self, n_rows, n_cols, seed, scale, device=None):
        self.shape = (int(n_rows), int(n_cols))
        self.seed, self.scale = int(seed), float(scale)
        self.device = _device(device)

    def sketch_into(self, A, b, out, row_offset=0):
        if row_offset % 4:
            raise ValueError("row shards must start at a multiple of 4 rows")
        K.sketch_gauss(A, self.shape[0], self.seed, self.scale, out, bvec=b, col_offset=row_offset)
        return out

    def apply(self, A):
        if A.shape[0] != self.shape[1]:
            raise ValueError(f"shape mismatch: {self.shape} @ {tuple(A.shape)}")
        out = torch.empty(self.shape[0], A.shape[1], dtype=F64, device=A.device)
        return self.sketch_into(A, None, out)

    def to_dense(self):
        return K.philox_normal_fill(self.shape[0], self.shape[1], self.seed, self.scale, device=self.device)

    def rmatvec(self, v, m_local=None, row_offset=0):
        m_local = self.shape[1] if m_local is None else m_local
        return K.gauss_rmatvec(self.shape[0], m_local, self.seed, self.scale, v, col_offset=row_offset)

    def column_slice(self, first, count):
        return self            # virtual: the column offset is passed to the kernel (sketch_into)


class SJLTOperator(SketchOperator):
    """d x m sparse sign operator, k nonzeros (+-1/sqrt(k)) per column, in index form."""

    def __init__(self, n_rows, rows, signs, validate=False):
        self.rows, self.signs = rows, signs
        self.shape = (int(n_rows), rows.shape[0])
        self.vec_nnz = rows.shape[1]
        self.scale = 1.0 / math.sqrt(self.vec_nnz)
        self.plan = K.SjltPlan(rows, signs, n_rows, validate=validate)

    def sketch_into(self, A, b, out, row_offset=0):
        n = A.shape[1]
        self.plan.apply(A, self.scale, out, bvec=b, out_b=None if b is None else out[:, n])
        return out

    def apply(self, A):
        if A.shape[0] != self.shape[1]:
            raise ValueError(f"shape mismatch: {self.shape} @ {tuple(A.shape)}")
        out = torch.empty(self.shape[0], A.shape[1], dtype=F64, device=A.device)
        return self.sketch_into(A, None, out)

    def column_slice(self, first, count):
        return SJLTOperator(self.shape[0], self.rows[first:first + count].contiguous(),
                            self.signs[first:first + count].contiguous())

    def rmatvec(self, v, m_local=None, row_offset=0):
        return K.sjlt_rmatvec(self.rows, self.signs, self.shape[0], v, self.scale)

    def to_dense(self):
        d, m = self.shape
        S = torch.zeros(d, m, dtype=F64, device=self.rows.device)
        cols = torch.arange(m, device=self.rows.device).repeat_interleave(self.vec_nnz)
        S.index_put_((self.rows.reshape(-1).long(), cols), self.signs.reshape(-1).to(F64) * self.scale,
                     accumulate=True)
        return S


class DenseOperator(SketchOperator):
    """Explicit dense operator on the device (replay of a reference ``S``; applied with the DMMA GEMM)."""

    def __init__(self, S):
        self.S = S
        self.shape = tuple(S.shape)

    def sketch_into(self, A, b, out, row_offset=0):
        n = A.shape[1]
        K.gemm(self.S, A, out=out[:, :n])
        if b is not None:
            out[:, n] = K.gemm(self.S, b.reshape(-1, 1)).reshape(-1)
        return out

    def apply(self, A):
        return K.gemm(self.S, A)

    def to_dense(self):
        return self.S

    def column_slice(self, first, count):
        return DenseOperator(self.S[:, first:first + count])

    def rmatvec(self, v, m_local=None, row_offset=0):
        return K.rmatvec(self.S, v)[:self.shape[1]].clone()


def as_device_operator(S, device=None):
    """Accept whatever a ``sketch_op_gen`` returned: one of our operators, a numpy ndarray, a torch
    tensor, or a scipy.sparse SJLT (fixed nnz per column, +-c values) as built by the reference."""
    if isinstance(S, SketchOperator):
        return S
    device = _device(device)
    if isinstance(S, torch.Tensor):
        return DenseOperator(S.to(device=device, dtype=F64))
    if isinstance(S, np.ndarray):
        return DenseOperator(torch.from_numpy(np.ascontiguousarray(S, dtype=np.float64)).to(device))
    try:
        import scipy.sparse as sps
    except ImportError:                                   # pragma: no cover
        sps = None
    if sps is not None and sps.issparse(S):
        C = sps.csc_matrix(S)
        counts = np.diff(C.indptr)
        k = int(counts[0]) if counts.size else 0
        mag = np.abs(C.data)
        if k > 0 and np.all(counts == k) and np.allclose(mag, mag[0], rtol=1e-14, atol=0) \
                and abs(mag[0] - 1.0 / math.sqrt(k)) < 1e-12:
            rows = torch.from_numpy(C.indices.reshape(-1, k).astype(np.int32)).to(device)
            signs = torch.from_numpy(np.sign(C.data).reshape(-1, k).astype(np.int8)).to(device)
            return SJLTOperator(C.shape[0], rows, signs, validate=True)
        return DenseOperator(torch.from_numpy(np.asarray(C.todense(), dtype=np.float64)).to(device))
    raise TypeError(f"unsupported sketching operator type {type(S)!r}")


def gaussian_operator(n_rows, n_cols, rng, normalize=True, device=None):
    """parla/utils/sketching.py:20-31.  N(0, 1/min(n_rows, n_cols)) entries (N(0,1) if not normalize)."""
    seed, _ = _draw_key(rng)
    scale = math.sqrt(1.0 / min(n_rows, n_cols)) if normalize else 1.0
    return GaussianOperator(n_rows, n_cols, seed, scale, device)


def sjlt_operator(n_rows, n_cols, rng, vec_nnz=8, device=None):
    """parla/utils/sketching.py:34-80.  Wide: vec_nnz distinct rows per column.  Tall: the transpose
    of the wide construction, returned as a dense device matrix (only used as a small test matrix)."""
    seed, rng = _draw_key(rng)
    device = _device(device)
    if n_cols >= n_rows:
        k = min(n_cols, vec_nnz)
        if n_rows < k:
            warnings.warn(f"Can't set {k} nonzeros per column for columns of length {n_rows}. "
                          "Sampling indices with replacement instead.")
        first, count = _SHARD if _SHARD is not None else (0, n_cols)
        rows, signs = K.sjlt_generate(n_rows, count, k, seed, col_offset=first, device=device)
        return SJLTOperator(n_rows, rows, signs)
    wide = sjlt_operator(n_cols, n_rows, np.random.default_rng(seed), vec_nnz, device)
    return DenseOperator(wide.to_dense().T.contiguous())
