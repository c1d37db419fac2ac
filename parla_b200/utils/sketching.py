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

    def sketch_into(self, A, b, out, row_offset=0, accumulate=False):
        if row_offset % 4:
            raise ValueError("row shards must start at a multiple of 4 rows")
        K.sketch_gauss(A, self.shape[0], self.seed, self.scale, out, bvec=b, col_offset=row_offset,
                       beta=1.0 if accumulate else 0.0)
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
        self._plan = K.SjltPlan(rows, signs, n_rows, validate=True) if validate else None

    @property
    def plan(self):
        """Destination-major plan, built on first use (an operator that is only sliced into row blocks, or only
        used through its adjoint, never needs the plan over all its columns)."""
        if self._plan is None:
            self._plan = K.SjltPlan(self.rows, self.signs, self.shape[0])
        return self._plan

    def sketch_into(self, A, b, out, row_offset=0, accumulate=False):
        n = A.shape[1]
        self.plan.apply(A, self.scale, out, bvec=b, out_b=None if b is None else out[:, n], accumulate=accumulate)
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


class SRCTOperator(SketchOperator):
    """Subsampled randomized cosine transform  S = R . DCT-II(ortho) . diag(e) . P  (d x m), the device
    form of what parla/utils/sketching.py:179-201 returns.  ``r`` (d sampled frequencies), ``e`` (m scaled
    signs) and ``perm`` (m) are the reference's ``sketch_data``.

    ``S @ A`` never forms the m x n transform: only the d sampled rows are evaluated, as a pruned two-level
    DCT (j = j1 + m1 j2) whose levels are FP64 tensor-core GEMMs -- 4 m n (m2 + d / m2) flops instead of the
    2 d m n of a dense product.  When m has no usable divisor, A is not contiguous, or the operator is applied
    to a row shard, the dense d x (rows) block of S is generated chunk by chunk and applied with the GEMM.
    """

    MAX_CHUNK_BYTES = 1 << 29
    STAGE_TIMINGS = None      # set to a dict to accumulate synchronised per-stage seconds (diagnostics only)

    def __init__(self, n_rows, n_cols, r, e, perm, device=None):
        self.device = _device(device)
        self.shape = (int(n_rows), int(n_cols))
        as_dev = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(
            device=self.device, dtype=dt)
        self.r, self.e = as_dev(r, torch.int64), as_dev(e, F64)
        d, m = self.shape
        self.perm = torch.arange(m, device=self.device) if perm is None else as_dev(perm, torch.int64)
        self.invperm = torch.empty_like(self.perm)
        self.invperm[self.perm] = torch.arange(m, device=self.device)
        self._plans = {}

    # ---- dense blocks of S (general path; also .T / to_dense for small operators)
    def dense_columns(self, first, count):
        """S[:, first:first+count]:  S[i, perm[j]] = e[j] c(r_i, j)."""
        return K.srct_weights(self.r, self.shape[1], first, count, jmap=self.invperm, e=self.e)

    def to_dense(self):
        return self.dense_columns(0, self.shape[1])

    def column_slice(self, first, count):
        return _SRCTSlice(self, first, count)

    def _apply_dense(self, A, b, out, first):
        d = self.shape[0]
        rows, n = A.shape
        step = max(2, min(rows + rows % 2, self.MAX_CHUNK_BYTES // (8 * d)))
        step -= step % 2
        sb = torch.zeros(d, 1, dtype=F64, device=self.device) if b is not None else None
        for t0 in range(0, rows, step):
            t1 = min(rows, t0 + step)
            Sd = self.dense_columns(first + t0, t1 - t0)
            K.gemm(Sd, A[t0:t1], beta=0.0 if t0 == 0 else 1.0, out=out[:, :n])
            if b is not None:
                K.gemm(Sd, b[t0:t1].reshape(-1, 1), beta=1.0, out=sb)
        if b is not None:
            out[:, n] = sb[:, 0]
        return out

    # ---- pruned two-level DCT
    def _plan(self, m2):
        if m2 in self._plans:
            return self._plans[m2]
        d, m = self.shape
        kap = self.r % (2 * m2)
        fold = torch.where(kap <= m2, kap, 2 * m2 - kap)
        sgn = torch.where(kap <= m2, 1.0, -1.0).to(F64)
        order = torch.argsort(fold, stable=True)
        fold_s = fold[order]
        vals, counts = torch.unique_consecutive(fold_s, return_counts=True)
        # level-1 table: rows [C_0, (C_1, S_1), ..., (C_{m2-1}, S_{m2-1}), C_{m2}] over j2 = 0 .. m2-1
        kk = np.arange(m2 + 1)[:, None] * np.arange(m2)[None, :] % (2 * m2)
        Cm, Sm = np.cos(np.pi * kk / m2), np.sin(np.pi * kk / m2)
        F = np.empty((2 * m2, m2))
        F[0], F[2 * m2 - 1] = Cm[0], Cm[m2]
        F[1:2 * m2 - 1:2], F[2:2 * m2 - 1:2] = Cm[1:m2], Sm[1:m2]
        plan = dict(order=order, k_sorted=self.r[order].contiguous(), sgn_sorted=sgn[order].contiguous(),
                    groups=list(zip(vals.tolist(), counts.tolist())),
                    F=torch.from_numpy(F).to(self.device))
        self._plans[m2] = plan
        return plan

    @staticmethod
    def choose_m2(m, d):
        """Divisor of m in [4, 512] minimising the flop count m2 + d / m2 (None: use the dense path)."""
        best = None
        for m2 in range(4, 513):
            if m % m2 == 0 and m // m2 >= 2:
                cost = m2 + d / m2
                if best is None or cost < best[0]:
                    best = (cost, m2)
        return None if best is None else best[1]

    def _apply_factored(self, A, b, out, m2):
        d, m = self.shape
        n = A.shape[1]
        m1 = m // m2
        plan = self._plan(m2)
        # memory this call may use: what the driver reports free + what torch's allocator holds unused
        free = torch.cuda.mem_get_info(self.device)[0] + torch.cuda.memory_reserved(self.device) \
            - torch.cuda.memory_allocated(self.device)
        w2_bytes = 16 * d * m1                               # all level-2 weights, kept across column blocks
        keep_w2 = w2_bytes <= free // 8
        budget = min((free - (w2_bytes if keep_w2 else 0)) // 2, 64 << 30)
        nb_max = max(2, int(budget // (24 * m)) // 2 * 2)    # Xp (m x w) + Y (2m x w) live together
        nblocks = -(-n // nb_max)
        nb = -(-n // nblocks)
        if -(-nb // 128) * 128 <= nb_max:
            nb = -(-nb // 128) * 128                         # whole GEMM tiles along N
        nb += nb % 2
        blocks = [(c0, min(nb, n - c0), False) for c0 in range(0, n, nb)]
        if b is not None:
            blocks.append((n, 1, True))                      # the right-hand side is its own (2-wide, padded) block
        Zs = torch.empty(d, n + (1 if b is not None else 0), dtype=F64, device=self.device)
        cache = {}
        rec = SRCTOperator.STAGE_TIMINGS

        def lap(name, t0):
            if rec is None:
                return 0.0
            import time
            torch.cuda.synchronize()
            now = time.time()
            rec[name] = rec.get(name, 0.0) + now - t0
            return now
        t0 = 0.0
        if rec is not None:
            import time
            torch.cuda.synchronize()
            t0 = time.time()

        def weights(idx, ks, sg, both):
            if idx not in cache:
                W2 = K.srct_weights(ks, m, 0, m1, sgn=sg if both else None, with_sin=both)
                if not keep_w2:
                    return W2
                cache[idx] = W2
            return cache[idx]

        for c0, width, is_rhs in blocks:
            w = width + width % 2                            # even leading dimension (16-byte GEMM loads)
            Xp = torch.empty(m, w, dtype=F64, device=self.device)
            if is_rhs:
                Xp[:, 0] = self.e * b[self.perm]
                Xp[:, 1] = 0.0
            else:
                if w != width:
                    Xp[:, width:] = 0.0
                K.gather_rows_scale(A, self.perm, self.e, c0, width, Xp)
            t0 = lap("gather", t0)
            Y = K.gemm(plan["F"], Xp.view(m2, m1 * w))       # level 1: (2 m2) x (m1 w)
            del Xp
            t0 = lap("level1", t0)
            s0 = 0
            for idx, (kappa, g) in enumerate(plan["groups"]):
                ks, sg = plan["k_sorted"][s0:s0 + g], plan["sgn_sorted"][s0:s0 + g]
                if kappa == 0 or kappa == m2:
                    row = 0 if kappa == 0 else 2 * m2 - 1
                    Z = K.gemm(weights(idx, ks, sg, False), Y[row].view(m1, w))
                else:
                    Z = K.gemm(weights(idx, ks, sg, True), Y[2 * kappa - 1:2 * kappa + 1].view(2 * m1, w))
                Zs[s0:s0 + g, c0:c0 + width] = Z[:, :width]
                s0 += g
            del Y
            t0 = lap("level2+weights", t0)
        out[:, :Zs.shape[1]].index_copy_(0, plan["order"], Zs)
        lap("scatter", t0)
        return out

    def sketch_into(self, A, b, out, row_offset=0, first=None):
        first = row_offset if first is None else first
        d, m = self.shape
        rows, n = A.shape
        whole = rows == m and first == 0
        m2 = self.choose_m2(m, d) if (whole and A.is_contiguous()) else None
        if m2 is None:
            return self._apply_dense(A, b, out, first)
        return self._apply_factored(A, b, out, m2)

    def apply(self, A):
        if A.shape[0] != self.shape[1]:
            raise ValueError(f"shape mismatch: {self.shape} @ {tuple(A.shape)}")
        out = torch.empty(self.shape[0], A.shape[1], dtype=F64, device=A.device)
        return self.sketch_into(A, None, out)

    def rmatvec(self, v, m_local=None, row_offset=0):
        """S[:, off:off+m_local]^T @ v  (`apply_srct(..., forward=False)` on one vector, sketching.py:162-175)."""
        d, m = self.shape
        m_local = m if m_local is None else m_local
        out = torch.empty(m_local, dtype=F64, device=self.device)
        step = max(2, min(K.PASS_MAX_N, self.MAX_CHUNK_BYTES // (8 * d)))
        for t0 in range(0, m_local, step):
            t1 = min(m_local, t0 + step)
            Sd = self.dense_columns(row_offset + t0, t1 - t0)
            out[t0:t1] = K.rmatvec(Sd, v)[:t1 - t0]
        return out


class _SRCTSlice(SketchOperator):
    """Columns [first, first+count) of an SRCTOperator (what a row shard of A meets)."""

    def __init__(self, base, first, count):
        self.base, self.first = base, int(first)
        self.shape = (base.shape[0], int(count))

    def sketch_into(self, A, b, out, row_offset=0):
        return self.base.sketch_into(A, b, out, first=self.first)

    def apply(self, A):
        out = torch.empty(self.shape[0], A.shape[1], dtype=F64, device=A.device)
        return self.sketch_into(A, None, out)

    def to_dense(self):
        return self.base.dense_columns(self.first, self.shape[1])

    def rmatvec(self, v, m_local=None, row_offset=0):
        return self.base.rmatvec(v, m_local=self.shape[1], row_offset=self.first)


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
    data = getattr(S, "sketch_data", None)                # the reference's SRCT LinearOperator (sketching.py:198)
    if data is not None and not getattr(S, "transposed", False):
        r, e, perm = data
        return SRCTOperator(S.shape[0], S.shape[1], r, e, perm, device)
    inner = getattr(S, "A", None) if data is None else S.T      # scipy's _TransposedLinearOperator / oracle .T
    if inner is not None and getattr(inner, "sketch_data", None) is not None:
        return DenseOperator(as_device_operator(inner, device).to_dense().T.contiguous())
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


def generate_srct(n_rows, n_cols, rng, device=None):
    """parla/utils/sketching.py:106-115: (r, e, perm) with r = small_dim distinct indices of the long axis,
    e = +-sqrt(big/small), perm a permutation of the long axis -- drawn on the device from one key of ``rng``."""
    seed, _ = _draw_key(rng)
    device = _device(device)
    big, small = max(n_rows, n_cols), min(n_rows, n_cols)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    r = torch.randperm(big, generator=gen, device=device)[:small].contiguous()
    e = torch.where(torch.rand(big, generator=gen, device=device, dtype=F64) > 0.5, 1.0, -1.0).to(F64)
    e = e * math.sqrt(big / small)
    perm = torch.randperm(big, generator=gen, device=device)
    return r, e, perm


def apply_srct(r, e, mat, perm=None, forward=True):
    """parla/utils/sketching.py:118-176 for device tensors (1-D or 2-D ``mat``)."""
    m = e.numel()
    S = SRCTOperator(r.numel(), m, r, e, perm, mat.device)
    if forward:
        return S @ mat
    if mat.dim() == 1:
        return S.rmatvec(mat)
    return K.gemm(S.to_dense(), mat, transa=True)


def srct_operator(n_rows, n_cols, rng, device=None):
    """parla/utils/sketching.py:179-201.  Wide: an SRCTOperator.  Tall: the transpose of the wide construction
    as a dense device matrix (only used as a small test matrix, e.g. by RS1)."""
    rng = np.random.default_rng(rng)
    r, e, perm = generate_srct(n_rows, n_cols, rng, device)
    if n_cols >= n_rows:
        return SRCTOperator(n_rows, n_cols, r, e, perm, device)
    wide = srct_operator(n_cols, n_rows, rng, device)
    return DenseOperator(wide.to_dense().T.contiguous())


def orthonormal_operator(n_rows, n_cols, rng, device=None):
    """parla/utils/sketching.py:9-17: the sign-normalised Q factor of a Gaussian matrix (orthonormal columns if
    tall, orthonormal rows if wide), as a dense device operator."""
    if n_rows < n_cols:
        return DenseOperator(orthonormal_operator(n_cols, n_rows, rng, device).S.T.contiguous())
    rng = np.random.default_rng(rng)
    G = gaussian_operator(n_rows, n_cols, rng, device=device).to_dense()
    Q, R = K.qr_economic(G)
    return DenseOperator((Q * torch.sign(torch.diagonal(R))).contiguous())


def sparse_sign_operator(n_rows, n_cols, rng, density=0.05, device=None):
    """parla/utils/sketching.py:83-103: iid sparse sign matrix, each entry nonzero with probability ``density``,
    values +-1/sqrt(min(n_rows, n_cols) * density).  Held densely on the device (applied with the DMMA GEMM)."""
    seed, _ = _draw_key(rng)
    device = _device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    for attempt in range(11):
        mask = torch.rand(n_rows, n_cols, generator=gen, device=device) < density
        if bool(mask.any()):
            break
        if attempt == 10:
            raise RuntimeError('Density too low.')
    sign = torch.where(torch.rand(n_rows, n_cols, generator=gen, device=device) < 0.5, -1.0, 1.0).to(F64)
    S = torch.where(mask, sign, torch.zeros((), dtype=F64, device=device)) / math.sqrt(min(n_rows, n_cols) * density)
    return DenseOperator(S.contiguous())


class SamplingOperator(SketchOperator):
    """Row-sampling operator: (S @ A)[i] = A[indices[i]]  (parla/utils/sketching.py:204-236, wide form)."""

    def __init__(self, n_rows, n_cols, indices):
        self.shape = (int(n_rows), int(n_cols))
        self.indices = indices

    def sketch_into(self, A, b, out, row_offset=0):
        n = A.shape[1]
        if A.shape[0] != self.shape[1]:
            raise ValueError("a sampling operator cannot be applied to a row shard")
        K.gather_rows_scale(A, self.indices, None, 0, n, out[:, :n])
        if b is not None:
            out[:, n] = b[self.indices]
        return out

    def apply(self, A):
        out = torch.empty(self.shape[0], A.shape[1], dtype=F64, device=A.device)
        return self.sketch_into(A, None, out)

    def to_dense(self):
        S = torch.zeros(self.shape, dtype=F64, device=self.indices.device)
        S[torch.arange(self.shape[0], device=self.indices.device), self.indices] = 1.0
        return S

    def rmatvec(self, v, m_local=None, row_offset=0):
        out = torch.zeros(self.shape[1], dtype=F64, device=v.device)
        out[self.indices] = v
        return out


def sampling_operator(n_rows, n_cols, rng, indices=None, device=None):
    """parla/utils/sketching.py:204-236: keep ``min(n_rows, n_cols)`` sorted, distinct positions of the long axis
    (drawn without replacement unless ``indices`` is given).  Wide: a SamplingOperator; tall: its transpose as a
    dense device matrix."""
    device = _device(device)
    pop, size = max(n_rows, n_cols), min(n_rows, n_cols)
    if indices is None:
        seed, _ = _draw_key(rng)
        gen = torch.Generator(device=device)
        gen.manual_seed(seed)
        idx = torch.sort(torch.randperm(pop, generator=gen, device=device)[:size])[0]
    else:
        idx = torch.as_tensor(np.asarray(indices), device=device).to(torch.int64)
        assert idx.numel() == size
    S = SamplingOperator(size, pop, idx.contiguous())
    if n_cols >= n_rows:
        return S
    return DenseOperator(S.to_dense().T.contiguous())
