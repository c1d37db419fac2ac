"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the PARLA randomized-sketching hot path.

A numpy/scipy *restatement* (not a copy) of the algorithms on the path named by
BASELINE.json:north_star.  Every function cites the reference file:line it follows
(paths relative to the reference checkout, BallisticLA/parla v0.1.4).

Pinning status: **pinned**.  The reference has no stored golden vectors, so the
oracle is pinned against outputs of the reference itself, imported in the build container:
``oracle/make_golden.py`` runs the reference and this oracle on identical inputs, asserts they
agree (x, residual, per-iteration error history, sketch operators bit-for-bit, low-rank
factors) and writes the fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py``
re-checks the oracle against those committed fixtures without needing the reference.

Restated here: the sketching operators (Gaussian, SJLT, SRCT), SPO / SSO1 / SPU1, LSQR, the lifted preconditioned
operator, PcSS2 (both branches), pcg / PcSS1, SPS1 (SVD and Nystrom preconditioners) / SPS2, RS1 / RF1 / QB1 / QB2 /
QB3, SVD1, EVD1 / EVD2, the interpolative / CUR decompositions, and the reference's test-problem generators.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import
this module.  The product path (``parla_b200``) never does.
"""
from __future__ import annotations

import math
import time
import warnings
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sps

EPS = float(np.finfo(np.float64).eps)


# ------------------------------------------------------------------------------------------
#  Sketching operators            (reference: parla/utils/sketching.py, comps/sketchers/oblivious.py)
# ------------------------------------------------------------------------------------------

def gaussian_operator(n_rows, n_cols, rng, normalize=True):
    """Dense iid Gaussian operator.  parla/utils/sketching.py:20-31.

    normalize=True draws N(0, 1/min(n_rows, n_cols)); otherwise N(0, 1).  The numpy
    Generator call sequence is the reference's, so the stream (and S) is identical.
    """
    rng = np.random.default_rng(rng)
    if not normalize:
        return rng.standard_normal((n_rows, n_cols))
    sd = np.sqrt(1.0 / min(n_rows, n_cols))
    return rng.normal(0.0, sd, (n_rows, n_cols))


def sjlt_operator(n_rows, n_cols, rng, vec_nnz=8):
    """Sparse Johnson-Lindenstrauss operator.  parla/utils/sketching.py:34-80.

    Wide case (n_cols >= n_rows): every column gets ``vec_nnz`` distinct row positions
    (one ``rng.choice`` per column, :63-65), then ONE ``rng.random(n_cols*vec_nnz)`` draw decides
    the signs (-1 where the uniform is <= 0.5, :69-70); values are +-1/sqrt(vec_nnz) (:71);
    result is CSC (:73-74).  Tall case: transpose of the wide construction, CSR (:78-79).
    """
    rng = np.random.default_rng(rng)
    if n_cols < n_rows:
        return sjlt_operator(n_cols, n_rows, rng, vec_nnz).T.tocsr()
    k = min(n_cols, vec_nnz)
    with_replacement = n_rows < k
    if with_replacement:
        warnings.warn(f"Can't set {k} nonzeros per column for columns of length {n_rows}. "
                      "Sampling indices with replacement instead.")
    picks = [rng.choice(n_rows, k, replace=with_replacement) for _ in range(n_cols)]
    ri = np.concatenate(picks)
    ci = np.repeat(np.arange(n_cols), k)
    sgn = np.ones(n_cols * k)
    sgn[rng.random(n_cols * k) <= 0.5] = -1.0
    sgn /= np.sqrt(k)
    return sps.coo_matrix((sgn, (ri, ci)), shape=(n_rows, n_cols)).tocsc()


class SkOpGA:
    """comps/sketchers/oblivious.py:38-45."""

    def __init__(self, normalize=True):
        self.normalize = normalize

    def __call__(self, n_rows, n_cols, rng):
        return gaussian_operator(n_rows, n_cols, rng, self.normalize)


class SkOpSJ:
    """comps/sketchers/oblivious.py:48-55."""

    def __init__(self, vec_nnz=8):
        self.vec_nnz = vec_nnz

    def __call__(self, n_rows, n_cols, rng):
        return sjlt_operator(n_rows, n_cols, rng, self.vec_nnz)


def generate_srct(n_rows, n_cols, rng):
    """Data of a subsampled randomized cosine transform, utils/sketching.py:106-115: the sampled
    rows ``r`` (small_dim distinct indices), the sign vector ``e`` scaled by sqrt(big/small), and a
    permutation of the long axis.  Generator call order: choice, random, permutation."""
    big, small = max(n_rows, n_cols), min(n_rows, n_cols)
    r = rng.choice(big, size=small, replace=False)
    u = rng.random(big)
    e = np.where(u > 0.5, 1.0, -1.0) * np.sqrt(big / small)
    perm = rng.permutation(big)
    return r, e, perm


def apply_srct(r, e, mat, perm=None, forward=True):
    """utils/sketching.py:118-176.  Forward: rows permuted, signs applied, orthonormal DCT-II along
    axis 0, rows ``r`` kept.  Adjoint: zero-fill, inverse DCT, signs, inverse permutation."""
    import scipy.fft as sfft
    if forward:
        x = mat if perm is None else mat[perm]
        x = x * (e[:, None] if mat.ndim > 1 else e)
        return sfft.dct(x, axis=0, norm='ortho')[r]
    full = np.zeros((e.size,) + mat.shape[1:])
    full[r] = mat
    full = sfft.idct(full, axis=0, norm='ortho')
    full = full * (e[:, None] if mat.ndim > 1 else e)
    if perm is not None:
        full = full[np.argsort(perm)]
    return full


class SrctOperator:
    """What utils/sketching.py:179-201 returns (there a scipy LinearOperator): wide d x m operator
    supporting ``S @ mat``, ``S.T @ mat``, ``.shape`` and carrying ``sketch_data = (r, e, perm)``."""

    def __init__(self, shape, data, transposed=False):
        self.shape, self.sketch_data, self.transposed = shape, data, transposed

    def __matmul__(self, mat):
        r, e, perm = self.sketch_data
        return apply_srct(r, e, np.asarray(mat), perm, forward=not self.transposed)

    @property
    def T(self):
        return SrctOperator(self.shape[::-1], self.sketch_data, not self.transposed)


def srct_operator(n_rows, n_cols, rng):
    """utils/sketching.py:179-201.  A tall request is the transpose of a wide construction; the
    reference draws the data once before branching (:186) and again inside the recursive call (:200),
    so the tall operator is built from the SECOND draw of the generator -- reproduced here."""
    data = generate_srct(n_rows, n_cols, rng)
    if n_cols >= n_rows:
        return SrctOperator((n_rows, n_cols), data)
    return srct_operator(n_cols, n_rows, rng).T


def srct_dense(r, e, perm, m):
    """The same operator as an explicit d x m matrix (definition of scipy's orthonormal DCT-II):
    S[i, perm[j]] = e[j] * f(r_i) * cos(pi (2j+1) r_i / (2m)),  f(0) = sqrt(1/m), f(k>0) = sqrt(2/m)."""
    j = np.arange(m)
    C = np.cos(np.pi * np.outer(r, 2 * j + 1) / (2 * m)) * np.where(r == 0, np.sqrt(1.0 / m), np.sqrt(2.0 / m))[:, None]
    S = np.zeros((r.size, m))
    S[:, perm if perm is not None else j] = C * e
    return S


class SkOpTC:
    """comps/sketchers/oblivious.py:68-72."""

    def __call__(self, n_rows, n_cols, rng):
        return srct_operator(n_rows, n_cols, rng)


def sjlt_index_form(S):
    """Fixed-nnz index form of a wide SJLT: (rows[int32, n_cols x k], signs[int8, n_cols x k], k).

    This is the layout the CUDA path consumes (SURVEY.md 8b "Injected S").  Requires every
    column to hold the same number of stored entries, which the construction guarantees when
    sampling without replacement.
    """
    S = sps.csc_matrix(S)
    counts = np.diff(S.indptr)
    k = int(counts[0])
    if not np.all(counts == k):
        raise ValueError("not a fixed-nnz-per-column operator")
    rows = S.indices.reshape(-1, k).astype(np.int32)
    signs = np.sign(S.data).reshape(-1, k).astype(np.int8)
    return rows, signs, k


# ------------------------------------------------------------------------------------------
#  Logging                         (reference: parla/comps/determiter/logging.py:4-94)
# ------------------------------------------------------------------------------------------

@dataclass
class SketchAndPrecondLog:
    time_sketch: float = 0.0
    time_factor: float = 0.0
    time_presolve: float = 0.0
    time_convert: float = 0.0
    time_iterate: float = 0.0
    times: np.ndarray | None = None
    errors: np.ndarray | None = None
    error_desc: str = "Fill in."

    @property
    def time_setup(self):
        # logging.py:63-70
        return self.time_sketch + self.time_factor + self.time_convert

    def wrap_up(self, iter_errors, init_error):
        # logging.py:72-94: amortised per-iteration clock, errors[0] = error at x = 0.
        iter_errors = np.atleast_1d(np.asarray(iter_errors, dtype=float))
        t0 = self.time_setup
        ramp = np.linspace(0.0, self.time_iterate, iter_errors.size, endpoint=True)
        self.times = np.concatenate(([t0], t0 + self.time_presolve + ramp))
        self.errors = np.concatenate(([init_error], iter_errors))


# ------------------------------------------------------------------------------------------
#  Preconditioned operator         (reference: parla/comps/preconditioning.py:16-79)
# ------------------------------------------------------------------------------------------

def svd_right_precond(A_ske):
    """preconditioning.py:70-79.  M = V / sigma restricted to the numerical rank."""
    U, sig, Vh = sla.svd(A_ske, full_matrices=False, check_finite=False)
    rank = int(np.count_nonzero(sig > sig[0] * A_ske.shape[1] * EPS))
    U, sig, Vh = U[:, :rank], sig[:rank], Vh[:rank, :]
    return Vh.T / sig, U, sig, Vh


class LiftedPrecondOperator:
    """[A; sqrt(delta) I] composed with a right preconditioner.  preconditioning.py:16-67.

    upper_tri=True : M = R^{-1} (triangular solves, :26-41)
    upper_tri=False: M = R      (dense products, :45-57)
    The reference materialises the lifted matrix (:6-13); here the identity block is applied
    implicitly, which is the same linear map.
    """

    def __init__(self, A, delta, R, upper_tri):
        self.A = A
        self.m, self.n = A.shape
        self.sd = math.sqrt(delta)
        self.R = R
        self.tri = bool(upper_tri)
        self.shape = (self.m + (self.n if delta > 0 else 0), R.shape[1])

    def precond(self, z):            # M_fwd, :40/:56
        if self.tri:
            return sla.solve_triangular(self.R, z, lower=False, check_finite=False)
        return self.R @ z

    def precond_t(self, w):          # M_adj, :41/:57
        if self.tri:
            return sla.solve_triangular(self.R, w, trans='T', lower=False, check_finite=False)
        return self.R.T @ w

    def matvec(self, z):             # forward(), :26-31 / :45-48
        x = self.precond(z)
        top = self.A @ x
        return top if self.sd == 0 else np.concatenate((top, self.sd * x))

    def rmatvec(self, u):            # adjoint(), :33-38 / :50-54
        w = self.A.T @ u[:self.m]
        if self.sd > 0:
            w = w + self.sd * u[self.m:]
        return self.precond_t(w)


# ------------------------------------------------------------------------------------------
#  LSQR                            (reference: parla/comps/determiter/lsqr.py:63-574)
# ------------------------------------------------------------------------------------------

def sym_ortho(a, b):
    """Stable Givens rotation, lsqr.py:63-95.  Returns (c, s, r)."""
    if b == 0:
        return np.sign(a), 0.0, abs(a)
    if a == 0:
        return 0.0, np.sign(b), abs(b)
    if abs(b) > abs(a):
        t = a / b
        s = np.sign(b) / math.sqrt(1.0 + t * t)
        return s * t, s, b / s
    t = b / a
    c = np.sign(a) / math.sqrt(1.0 + t * t)
    return c, c * t, a / c


@dataclass
class LsqrResult:
    x: np.ndarray
    istop: int
    itn: int
    r1norm: float
    r2norm: float
    anorm: float
    acond: float
    arnorms: np.ndarray      # history; a 0-d value on the early-exit path (lsqr.py:392-395)
    xnorm: float


def lsqr(op, b, atol, btol, iter_lim, conlim=1e8, x0=None):
    """Paige-Saunders LSQR with PARLA's modifications (damp = 0 always on this path).

    lsqr.py:98-574.  Differences of the reference from SciPy that matter and are kept:
      * arnorm history: arnorms[itn] is recorded BEFORE step itn+1 (:413), trimmed at :572;
      * x0 is used directly: x = x0, u = b - A x0 (:367-370);
      * early exit when alfa*beta == 0 returns the scalar arnorm (:392-395).
    Stopping rules :500-526 (istop 1..7).
    """
    b = np.asarray(b, dtype=float).reshape(-1)
    n = op.shape[1]
    ctol = 1.0 / conlim if conlim > 0 else 0.0
    anorm = acond = ddnorm = xnorm = xxnorm = z = 0.0
    cs2, sn2 = -1.0, 0.0
    bnorm = float(np.linalg.norm(b))
    if x0 is None:
        x = np.zeros(n)
        u = b.copy()
        beta = bnorm
    else:
        x = np.array(x0, dtype=float)
        u = b - op.matvec(x)
        beta = float(np.linalg.norm(u))
    if beta > 0:
        u = u / beta
        v = op.rmatvec(u)
        alfa = float(np.linalg.norm(v))
    else:
        v = x.copy()
        alfa = 0.0
    if alfa > 0:
        v = v / alfa
    w = v.copy()
    rhobar, phibar = alfa, beta
    rnorm = beta
    arnorm = alfa * beta
    if arnorm == 0:
        return LsqrResult(x, 0, 0, rnorm, rnorm, anorm, acond, np.float64(arnorm), xnorm)

    hist = -np.ones(iter_lim)
    itn, istop = 0, 0
    while itn < iter_lim:
        hist[itn] = arnorm
        itn += 1
        # bidiagonalisation step (:421-430)
        u = op.matvec(v) - alfa * u
        beta = float(np.linalg.norm(u))
        if beta > 0:
            u = u / beta
            anorm = math.sqrt(anorm * anorm + alfa * alfa + beta * beta)
            v = op.rmatvec(u) - beta * v
            alfa = float(np.linalg.norm(v))
            if alfa > 0:
                v = v / alfa
        # damp == 0  =>  rhobar1 = rhobar, psi = 0 (:434-438)
        cs, sn, rho = sym_ortho(rhobar, beta)
        theta = sn * alfa
        rhobar = -cs * alfa
        phi = cs * phibar
        phibar = sn * phibar
        tau = sn * phi
        # solution / search-direction update (:449-457)
        dk_sq = float(np.dot(w, w)) / (rho * rho)
        x = x + (phi / rho) * w
        w = v + (-theta / rho) * w
        ddnorm += dk_sq
        # norm(x) estimate via a second rotation (:462-471)
        delta_ = sn2 * rho
        gambar = -cs2 * rho
        rhs = phi - delta_ * z
        zbar = rhs / gambar
        xnorm = math.sqrt(xxnorm + zbar * zbar)
        gamma = math.sqrt(gambar * gambar + theta * theta)
        cs2, sn2 = gambar / gamma, theta / gamma
        z = rhs / gamma
        xxnorm += z * z
        # convergence estimates (:476-498)
        acond = anorm * math.sqrt(ddnorm)
        rnorm = math.sqrt(phibar * phibar)
        arnorm = alfa * abs(tau)
        test1 = rnorm / bnorm
        test2 = arnorm / (anorm * rnorm + EPS)
        test3 = 1.0 / (acond + EPS)
        t1 = test1 / (1.0 + anorm * xnorm / bnorm)
        rtol = btol + atol * anorm * xnorm / bnorm
        # stopping rules, later assignments win (:504-526)
        if itn >= iter_lim:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + t1 <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if istop:
            break
    return LsqrResult(x, istop, itn, rnorm, rnorm, anorm, acond, hist[hist > -1], xnorm)


# ------------------------------------------------------------------------------------------
#  PcSS2 + SPO / SSO1              (reference: determiter/saddle.py:180-217, drivers/least_squares.py)
# ------------------------------------------------------------------------------------------

def pcss2_overdetermined(A, b, delta, tol, iter_lim, R, upper_tri, z0):
    """Over-determined branch of PcSS2.__call__, saddle.py:187-201.  Returns (x, y, arnorms)."""
    m, n = A.shape
    op = LiftedPrecondOperator(A, delta, R, upper_tri)
    rhs = np.concatenate((b, np.zeros(n))) if delta > 0 else b
    res = lsqr(op, rhs, atol=tol, btol=tol, iter_lim=iter_lim, x0=z0)
    x = op.precond(res.x)
    y = rhs[:m] - A @ x
    return x, y, res.arnorms


class _Transposed:
    """The adjoint view `A_pc.T` that saddle.py:206 hands to lsqr."""

    def __init__(self, op):
        self.op = op
        self.shape = (op.shape[1], op.shape[0])

    def matvec(self, y):
        return self.op.rmatvec(y)

    def rmatvec(self, z):
        return self.op.matvec(z)


def pcss2_underdetermined(A, c, delta, tol, iter_lim, R, upper_tri):
    """Under-determined branch of PcSS2.__call__, saddle.py:203-214.  Returns (x, y, arnorms):
    y is the minimum-norm solution of (A M)' y = M' c; x = (A'y - c)/delta when delta > 0, else NaN."""
    m, n = A.shape
    op = LiftedPrecondOperator(A, delta, R, upper_tri)
    c_pc = op.precond_t(c)
    res = lsqr(_Transposed(op), c_pc, atol=tol, btol=tol, iter_lim=iter_lim)
    y = res.x
    if delta > 0:
        y = y[:m]
        x = (A.T @ y - c) / delta
    else:
        x = np.nan * np.empty(n)
    return x, y, res.arnorms


class SPU1:
    """SVD-based sketch-and-precondition for under-determined least squares
    (min ||y|| s.t. A'y = c), least_squares.py:425-494."""

    def __init__(self, sketch_op_gen, sampling_factor):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor

    def __call__(self, A, c, tol, iter_lim, rng, logging=True):
        m, n = A.shape
        d = dim_checks(self.sampling_factor, m, n)
        rng = np.random.default_rng(rng)
        clock = time.time if logging else (lambda: 0.0)
        log = SketchAndPrecondLog()
        t = clock()
        S = self.sketch_op_gen(d, m, rng)                       # :469-471
        A_ske = S @ A
        log.time_sketch = clock() - t
        t = clock()
        M, _U, _s, _Vh = svd_right_precond(A_ske)               # :475
        log.time_factor = clock() - t
        t = clock()
        x, y, arnorms = pcss2_underdetermined(A, c, 0.0, tol, iter_lim, M, False)    # :480
        log.time_iterate = clock() - t
        if logging:
            log.wrap_up(arnorms, float(np.linalg.norm(A @ (M @ (M.T @ c)))))        # :486
        return y, log


def dim_checks(sampling_factor, n_rows, n_cols):
    """least_squares.py:89-104."""
    assert n_rows >= n_cols
    d = int(sampling_factor * n_cols)
    if d > n_rows:
        warnings.warn(f"embedding dimension d={d} exceeds the {n_rows} rows of the data matrix; "
                      f"proceeding with d={n_rows} (very inefficient).")
        d = n_rows
    assert d >= n_cols
    return d


class SPO:
    """Sketch-and-precondition least squares, least_squares.py:206-369 (mode qr/chol/svd).

    "SAP1" == mode 'qr', "SAP2" == mode 'svd' (SURVEY.md section 0).
    """

    def __init__(self, sketch_op_gen, sampling_factor, mode='qr'):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.mode = mode

    def __call__(self, A, b, delta, tol, iter_lim, rng, logging=True):
        m, n = A.shape
        sd = math.sqrt(delta)
        d = dim_checks(self.sampling_factor, m, n)
        rng = np.random.default_rng(rng)
        clock = time.time if logging else (lambda: 0.0)
        log = SketchAndPrecondLog()

        t = clock()
        S = self.sketch_op_gen(d, m, rng)                       # :302
        A_ske = S @ A                                           # :303
        log.time_sketch = clock() - t

        t = clock()
        if self.mode == 'qr':                                   # :306-316
            if delta > 0:
                A_ske = np.vstack((A_ske, sd * np.eye(n)))
            Q, R = sla.qr(A_ske, mode='economic')
            log.time_factor = clock() - t
            t = clock()
            b_ske = S @ b
            z_ske = Q[:d].T @ b_ske
            x_ske = sla.solve_triangular(R, z_ske, lower=False)
        elif self.mode == 'chol':                               # :317-329
            G = A_ske.T @ A_ske + delta * np.eye(n)
            R = sla.cholesky(G, lower=False, check_finite=False)
            log.time_factor = clock() - t
            t = clock()
            b_ske = S @ b
            z_ske = sla.solve_triangular(R, A_ske.T @ b_ske, lower=False, trans='T')
            x_ske = sla.solve_triangular(R, z_ske, lower=False)
        elif self.mode == 'svd':                                # :330-339
            if delta > 0:
                A_ske = np.vstack((A_ske, sd * np.eye(n)))
            R, U, _sig, _Vh = svd_right_precond(A_ske)
            log.time_factor = clock() - t
            t = clock()
            b_ske = S @ b
            z_ske = U[:d].T @ b_ske
            x_ske = R @ z_ske
        else:
            raise ValueError()                                  # :341

        # presolve acceptance test (:344-352)
        r = A @ x_ske - b
        if delta > 0:
            r = np.concatenate((r, sd * x_ske))
        rel_err = np.linalg.norm(r) / np.linalg.norm(b)
        if rel_err >= 1 or (rel_err > 1e-15 and R.shape[0] != R.shape[1]):
            z_ske = None
        log.time_presolve = clock() - t

        t = clock()
        tri = self.mode in ('qr', 'chol')
        x, y, arnorms = pcss2_overdetermined(A, b, delta, tol, iter_lim, R, tri, z_ske)   # :356-357
        log.time_iterate = clock() - t

        if logging:                                             # :360-367
            g = A.T @ b
            g = sla.solve_triangular(R, g, trans='T', lower=False) if tri else R.T @ g
            log.wrap_up(arnorms, float(np.linalg.norm(g)))
            log.error_desc = "2-norm of the residual of the preconditioned normal equations"
        self.last_y = y
        return x, log


# ------------------------------------------------------------------------------------------
#  Interpolative decompositions    (reference: comps/interpolative.py, drivers/interpolative.py)
# ------------------------------------------------------------------------------------------

def qrcp_osid(Y, k, axis):
    """Rank-k one-sided ID of Y from a column-pivoted QR, comps/interpolative.py:11-67.
    axis=1: (X, Js) with Y ~ Y[:, Js] X, X[:, Js] = I.  axis=0: (Z, Is) with Y ~ Z Y[Is, :]."""
    if axis == 0:
        X, Is = qrcp_osid(Y.T, k, 1)
        return X.T, Is
    if axis != 1:
        raise ValueError()
    _, R, J = sla.qr(Y, mode='economic', pivoting=True)
    T = sla.solve_triangular(R[:k, :k], R[:k, k:], lower=False)
    X = np.zeros((k, Y.shape[1]))
    X[:, J] = np.hstack((np.eye(k), T))
    return X, J[:k]


def _pivots(Y, k):
    return sla.qr(Y, mode='economic', pivoting=True)[2][:k]


class ROCS1:
    """Row / column selection: QRCP pivots of a sketch, comps/interpolative.py:131-157."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, over, axis, rng):
        rng = np.random.default_rng(rng)
        if axis == 0:
            return _pivots((A @ self.sk_op(A, k + over, rng)).T, k)
        if axis == 1:
            return _pivots(self.sk_op(A.T, k + over, rng).T @ A, k)
        raise ValueError()


class OSID1:
    """Sketch, then ID of the sketch, drivers/interpolative.py:97-134."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, over, axis, rng):
        rng = np.random.default_rng(rng)
        if axis == 0:
            return qrcp_osid(A @ self.sk_op(A, k + over, rng), k, 0)
        if axis == 1:
            return qrcp_osid(self.sk_op(A.T, k + over, rng).T @ A, k, 1)
        raise ValueError()


class OSID2:
    """Skeleton from the QRCP of a sketch, coefficients by a pseudo-inverse, drivers/interpolative.py:146-176."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, over, axis, rng):
        rng = np.random.default_rng(rng)
        if axis == 0:
            Is = _pivots((A @ self.sk_op(A, k + over, rng)).T, k)
            return sla.lstsq(A[Is, :].T, A.T)[0].T, Is              # A pinv(A[Is, :])
        if axis == 1:
            Js = _pivots(self.sk_op(A.T, k + over, rng).T @ A, k)
            return sla.lstsq(A[:, Js], A)[0], Js                    # pinv(A[:, Js]) A
        raise ValueError()


class TSID1:
    """Two-sided ID A ~ Z A[Is, Js] X, drivers/interpolative.py:257-278."""

    def __init__(self, osid):
        self.osid = osid

    def __call__(self, A, k, over, rng):
        rng = np.random.default_rng(rng)
        if A.shape[0] > A.shape[1]:
            X, Js = self.osid(A, k, over, axis=1, rng=rng)
            Z, Is = qrcp_osid(A[:, Js], k, 0)
        else:
            Z, Is = self.osid(A, k, over, axis=0, rng=rng)
            X, Js = qrcp_osid(A[Is, :], k, 1)
        return Z, Is, X, Js


class CUR1:
    """CUR decomposition A ~ A[:, Js] U A[Is, :], drivers/interpolative.py:358-387."""

    def __init__(self, osid):
        self.osid = osid

    def __call__(self, A, k, over, rng):
        rng = np.random.default_rng(rng)
        if A.shape[0] > A.shape[1]:
            X, Js = self.osid(A, k, over, axis=1, rng=rng)
            Is = _pivots(A[:, Js].T, k)
            U = sla.lstsq(A[Is, :].T, X.T)[0].T                     # X pinv(A[Is, :])
        else:
            Z, Is = self.osid(A, k, over, axis=0, rng=rng)
            Js = _pivots(A[Is, :], k)
            U = sla.lstsq(A[:, Js], Z)[0]                           # pinv(A[:, Js]) Z
        return Js, U, Is


# ------------------------------------------------------------------------------------------
#  Saddle-point systems            (reference: comps/determiter/pcg.py, comps/determiter/saddle.py:88-176,
#                                   drivers/saddlesys.py:89-327)
# ------------------------------------------------------------------------------------------

def pcg(mv_mat, rhs, mv_pre, iter_lim, tol, x0):
    """Preconditioned conjugate gradients, pcg.py:5-47.

    History entry i is ||rhs - mat x_i||_2 at the START of iteration i (:28); the residual is
    recomputed from x on iterations 0, 10, 20, ... (:34-35) and updated recursively otherwise
    (:37); the loop runs while ||r|| > tol * ||r_0|| (:23,:26).  Returns (x, history[:iters]).
    """
    x = np.array(x0, dtype=np.float64, copy=True)
    r = rhs - mv_mat(x)
    p = mv_pre(r)
    rz = float(r @ p)
    err = float(np.linalg.norm(r))
    stop_at = tol * err
    hist = []
    it = 0
    while it < iter_lim and err > stop_at:
        hist.append(err)
        q = mv_mat(p)
        step = rz / float(p @ q)
        x = x + step * p
        r = rhs - mv_mat(x) if it % 10 == 0 else r - step * q
        err = float(np.linalg.norm(r))
        s = mv_pre(r)
        rz_prev, rz = rz, float(r @ s)
        p = s + (rz / rz_prev) * p
        it += 1
    return x, np.array(hist, dtype=np.float64)


def pcss1(A, b, c, delta, tol, iter_lim, R, upper_tri, z0):
    """PcSS1.__call__, saddle.py:96-176: PCG on (A'A + delta I) x = A'b - c with the
    preconditioner R R' (full rank) or R~ R~' + (I - V V') (low rank, :121-131,:134-142).
    Returns (x, y = b - A x, history).  R is not modified (the reference rescales it in place)."""
    m, n = A.shape
    if b is None:
        b = np.zeros(m)
    if b.ndim != 1:
        raise NotImplementedError()
    if upper_tri:
        raise NotImplementedError()                             # :117-118
    full = R.shape[1] == n
    if full:
        Rs, V = R, None
    else:
        sv = 1.0 / np.linalg.norm(R, axis=0)                    # :124
        V = R * sv
        sv = np.sqrt(sv ** 2 + delta)                           # :127-129
        Rs = V / (sv / sv[-1])                                  # :130-131

    def mv_pre(vec):                                            # :134-142
        out = Rs @ (Rs.T @ vec)
        if not full:
            out = out + vec - V @ (V.T @ vec)
        return out

    def mv_gram(vec):                                           # :144-148
        return A.T @ (A @ vec) + delta * vec

    rhs = A.T @ b
    if c is not None:
        rhs = rhs - c
    x0 = np.zeros(n) if (z0 is None or not full) else Rs @ z0   # :154-158
    x, hist = pcg(mv_gram, rhs, mv_pre, iter_lim, tol, x0)
    return x, b - A @ x, hist


class SPS1:
    """SVD / Nystrom sketch-and-precondition for saddle systems with PCG, saddlesys.py:89-223."""

    def __init__(self, sketch_op_gen, sampling_factor, iterative_solver=None):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.iterative_solver = iterative_solver if iterative_solver is not None else pcss1
        self.nystrom_strategy = 'left'

    def __call__(self, A, b, c, delta, tol, iter_lim, rng, logging=True):
        m, n = A.shape
        d = int(self.sampling_factor * n)                       # :134 (no dim_checks here)
        rng = np.random.default_rng(rng)
        assert self.nystrom_strategy in ('left', 'right')
        if b is None:
            b = np.zeros(m)
        clock = time.time if logging else (lambda: 0.0)
        log = SketchAndPrecondLog()
        if d >= n:                                              # :146-157
            t = clock()
            S = self.sketch_op_gen(d, m, rng)
            A_ske = S @ A
            if delta > 0:
                A_ske = np.vstack((A_ske, math.sqrt(delta) * np.eye(n)))
            log.time_sketch = clock() - t
            t = clock()
            M, _U, sigma, Vh = svd_right_precond(A_ske)
            log.time_factor = clock() - t
        elif self.nystrom_strategy == 'right':                  # :159-170
            t = clock()
            S = self.sketch_op_gen(n, d, rng)
            Y = A @ S
            log.time_sketch = clock() - t
            t = clock()
            Q = orth(Y)
            _U, sigma, Vh = sla.svd(Q.T @ A, full_matrices=False)
            M = Vh.T / sigma
            log.time_factor = clock() - t
        else:                                                   # :171-184
            t = clock()
            S = self.sketch_op_gen(d, m, rng)
            A_ske = S @ A
            log.time_sketch = clock() - t
            t = clock()
            V = orth(A_ske.T)
            _U, sigma, Wt = sla.svd(A @ V, full_matrices=False)
            M = V @ (Wt.T / sigma)
            log.time_factor = clock() - t

        rhs = A.T @ b
        if c is not None:
            rhs = rhs - c
        t = clock()
        z_ske = None
        if d >= n:                                              # :198-204
            z_ske = (Vh @ rhs) / sigma
            x_ske = M @ z_ske
            rhs_pc = M.T @ rhs
            lhs_pc = M.T @ (A.T @ (A @ x_ske) + delta * x_ske)
            if np.linalg.norm(lhs_pc - rhs_pc) >= np.linalg.norm(rhs_pc):
                z_ske = None
        log.time_presolve = clock() - t

        t = clock()
        x, y, hist = self.iterative_solver(A, b, c, delta, tol, iter_lim, M, False, z_ske)   # :212
        log.time_iterate = clock() - t
        if logging:
            log.wrap_up(hist, float(np.linalg.norm(rhs)))       # :219
        return x, y, log


class SPS2:
    """Sketch, reduce the saddle system to over-determined least squares, precondition, LSQR.
    saddlesys.py:226-327."""

    def __init__(self, sketch_op_gen, sampling_factor, iterative_solver=None):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.iterative_solver = iterative_solver

    def __call__(self, A, b, c, delta, tol, iter_lim, rng, logging=False):
        m, n = A.shape
        sd = math.sqrt(delta)
        d = dim_checks(self.sampling_factor, m, n)
        rng = np.random.default_rng(rng)
        if b is None:
            b = np.zeros(m)
        clock = time.time if logging else (lambda: 0.0)
        log = SketchAndPrecondLog()

        t = clock()                                             # :275-279
        S = self.sketch_op_gen(d, m, rng)
        A_ske = S @ A
        if delta > 0:
            A_ske = np.vstack((A_ske, sd * np.eye(n)))
        log.time_sketch = clock() - t

        t = clock()                                             # :282-284
        M, U, sigma, Vh = svd_right_precond(A_ske)
        log.time_factor = clock() - t

        t = clock()                                             # :287-295
        A_aug = np.vstack((A, sd * np.eye(n))) if delta > 0 else A
        b_aug = np.concatenate((b, np.zeros(n))) if delta > 0 else b.copy()
        if c is not None and np.linalg.norm(c) > 0:
            v = U @ ((Vh @ c) / sigma)
            b_aug[:m] -= S.T @ v[:d]
            if delta > 0:
                b_aug[m:] -= v[d:]
        log.time_convert = clock() - t

        t = clock()                                             # :298-304
        z_ske = U[:d].T @ (S @ b_aug[:m])
        if delta > 0:
            z_ske = z_ske + U[d:].T @ b_aug[m:]
        x_ske = M @ z_ske
        if np.linalg.norm(b_aug - A_aug @ x_ske) >= np.linalg.norm(b_aug):
            z_ske = None
        log.time_presolve = clock() - t

        t = clock()                                             # :307-313
        solver = self.iterative_solver if self.iterative_solver is not None else pcss2_overdetermined_full
        x, _y_aug, hist = solver(A_aug, b_aug, None, 0.0, tol, iter_lim, M, False, z_ske)
        log.time_iterate = clock() - t
        y = b - A @ x
        if logging:                                             # :316-321
            log.wrap_up(hist, float(np.linalg.norm(M.T @ (A_aug.T @ b_aug))))
        return x, y, log


def pcss2_overdetermined_full(A, b, c, delta, tol, iter_lim, R, upper_tri, z0):
    """PcSS2 with the PrecondSaddleSolver signature (saddle.py:187), over-determined branch only."""
    assert c is None or np.linalg.norm(c) == 0
    return pcss2_overdetermined(A, b, delta, tol, iter_lim, R, upper_tri, z0)


class SSO1:
    """Sketch-and-solve, least_squares.py:114-189."""

    def __init__(self, sketch_op_gen, sampling_factor, lapack_driver=None, overwrite_sketch=True):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.lapack_driver = lapack_driver

    def __call__(self, A, b, delta, tol, iter_lim, rng, logging=True):
        if not np.isnan(tol):
            warnings.warn('SSO1 cannot control approximation error; "tol" is ignored.')
        if iter_lim > 1:
            warnings.warn('SSO1 is not iterative; "iter_lim" is ignored.')
        m, n = A.shape
        d = dim_checks(self.sampling_factor, m, n)
        rng = np.random.default_rng(rng)
        clock = time.time if logging else (lambda: 0.0)
        t = clock()
        S = self.sketch_op_gen(d, m, rng)
        A_ske, b_ske = S @ A, S @ b
        log = {'time_sketch': clock() - t}
        t = clock()
        if delta > 0:
            A_ske = np.vstack((A_ske, math.sqrt(delta) * np.eye(n)))
            b_ske = np.concatenate((b_ske, np.zeros(n)))
        x = sla.lstsq(A_ske, b_ske, cond=None, check_finite=False,
                      lapack_driver=self.lapack_driver)[0]
        log['time_solve'] = clock() - t
        return x, log


# ------------------------------------------------------------------------------------------
#  Low-rank path: orth / RS1 / RF1 / QB1 / QB2 / SVD1 / EVD1
# ------------------------------------------------------------------------------------------

def orth(S):
    """utils/linalg_wrappers.py:6-7: Q factor of an economic Householder QR."""
    return sla.qr(S, mode='economic')[0]


class RS1:
    """Power-iteration row sketcher, comps/sketchers/aware.py:118-184."""

    def __init__(self, sketch_op_gen, num_pass, stabilizer, passes_per_stab):
        self.sketch_op_gen = sketch_op_gen
        self.num_pass = num_pass
        self.stabilizer = stabilizer
        self.passes_per_stab = passes_per_stab

    def __call__(self, A, k, rng):
        assert self.num_pass >= 0                               # :160
        rng = np.random.default_rng(rng)
        done = 0
        if self.num_pass % 2 == 0:                              # :163-164
            S = self.sketch_op_gen(A.shape[1], k, rng)
        else:                                                   # :165-169
            S = A.T @ self.sketch_op_gen(A.shape[0], k, rng)
            done = 1
            if self.passes_per_stab == 1:
                S = self.stabilizer(S)
        for _ in range((self.num_pass - done) // 2):            # :170-183
            S = A @ S
            done += 1
            if done % self.passes_per_stab == 0:
                S = self.stabilizer(S)
            S = A.T @ S
            done += 1
            if done % self.passes_per_stab == 0:
                S = self.stabilizer(S)
        return S


class RF1:
    """Rangefinder, comps/rangefinders.py:126-188."""

    def __init__(self, rso):
        self.rso = rso

    def __call__(self, A, k, tol, rng):
        assert 0 < k <= min(A.shape)                            # :176-177
        if not np.isnan(tol):
            warnings.warn('RF1 cannot control approximation error; "tol" is ignored.')
        rng = np.random.default_rng(rng)
        S = self.rso(A, k, rng)
        return sla.qr(A @ S, mode='economic')[0]                # :186-187


class QB1:
    """comps/qb.py:286-352."""

    def __init__(self, rf):
        self.rangefinder = rf

    def __call__(self, A, k, tol, rng):
        assert 0 < k <= min(A.shape)
        if not np.isnan(tol):
            assert 0 <= tol < 1
        rng = np.random.default_rng(rng)
        Q = self.rangefinder(A, k, tol, rng)
        return Q, Q.T @ A


class QB2:
    """Blocked, tolerance-controlled QB with in-place deflation, comps/qb.py:355-482."""

    def __init__(self, rf, blk, overwrite_a):
        self.rangefinder = rf
        self.blk = blk
        self.overwrite_a = overwrite_a

    def __call__(self, A, k, tol, rng):
        if not self.overwrite_a:                                # :442-443
            A = np.array(A, copy=True)
        assert k > 0
        lim = min(A.shape)
        if k > lim:                                             # :445-452
            warnings.warn(f"target rank k={k} exceeds min{A.shape}; using k={lim}.")
            k = lim
        use_tol = (not np.isnan(tol)) and tol > 0               # :454
        if use_tol:
            sq_norm = np.linalg.norm(A, 'fro') ** 2
            stop_at = sq_norm * tol ** 2
        rng = np.random.default_rng(rng)
        m, n = A.shape
        Q = np.empty((m, 0))
        B = np.empty((0, n))
        blk = self.blk
        while True:                                             # :463-481
            if B.shape[0] + blk > k:
                blk = k - B.shape[0]
            Qi = self.rangefinder(A, blk, np.nan, rng)
            Qi = Qi - Q @ (Q.T @ Qi)                            # project_out, :603-613
            Qi = sla.qr(Qi, mode='economic')[0]
            Bi = Qi.T @ A
            Q = np.column_stack((Q, Qi))
            B = np.vstack((B, Bi))
            A -= Qi @ Bi
            if use_tol:
                sq_norm -= np.linalg.norm(Bi, 'fro') ** 2
                if sq_norm <= stop_at:
                    break
            if B.shape[0] >= k:
                break
        return Q, B


class QB3:
    """Single-pass-style blocked QB from the sketches G = A S and H = A'G, comps/qb.py:484-598
    ([YGL:2018, Algorithm 4] with an arbitrary row sketcher).  tol > 0 enables early stopping."""

    def __init__(self, sk_op, blk):
        self.sk_op = sk_op
        self.blk = blk

    def __call__(self, A, k, tol, rng):
        assert k > 0 and k < min(A.shape)                         # :555-556
        use_tol = (not np.isnan(tol)) and tol > 0
        if use_tol:
            sq_norm = np.linalg.norm(A, 'fro') ** 2
            stop_at = sq_norm * tol ** 2
        rng = np.random.default_rng(rng)
        S = self.sk_op(A, k, rng)
        if not isinstance(S, np.ndarray):
            raise RuntimeError(f"QB3 needs a dense sketching matrix, got {type(S)}")      # :568-574
        G = A @ S
        H = A.T @ G
        m, n = A.shape
        Q, B = np.empty((m, 0)), np.empty((0, n))
        for lo in range(0, k, self.blk):                          # :577-597
            hi = min(lo + self.blk, k)
            BS = B @ S[:, lo:hi]
            Y = G[:, lo:hi] - Q @ BS
            Qi, Ri = sla.qr(Y, mode='economic')
            Qi = Qi - Q @ (Q.T @ Qi)
            Qi, Rhat = sla.qr(Qi, mode='economic')
            Ri = Rhat @ Ri
            Bi = H[:, lo:hi].T - (Y.T @ Q) @ B - BS.T @ B
            Bi = sla.solve_triangular(Ri, Bi, trans='T', lower=False)
            Q, B = np.hstack((Q, Qi)), np.vstack((B, Bi))
            if use_tol:
                sq_norm -= np.linalg.norm(Bi, 'fro') ** 2
                if sq_norm <= stop_at:
                    break
        return Q, B


class EVD2:
    """Fixed-rank PSD eigendecomposition from a regularised Nystrom approximation, drivers/evd.py:290-381
    (Tropp, Yurtsever, Udell, Cevher 2017, Algorithm 3)."""

    def __init__(self, sk_op):
        self.sk_op = sk_op

    def __call__(self, A, k, tol, over, rng):
        n = A.shape[0]
        assert k > 0 and k < n                                    # :352-354
        if not np.isnan(tol):
            warnings.warn("EVD2 cannot control the approximation error; 'tol' is ignored.")
        rng = np.random.default_rng(rng)
        S = self.sk_op(A, k + over, rng)
        Y = A @ S
        nu = np.sqrt(n) * EPS * np.linalg.norm(Y)                 # :365 temporary regularisation
        Y = Y + nu * S
        R = sla.cholesky(S.T @ Y, lower=False)
        Bm = sla.solve_triangular(R, Y.T, lower=False, trans='T').T
        V, sigma, _ = sla.svd(Bm, full_matrices=False)
        r = min([k] + [i for i in range(k - 1) if sigma[i + 1] ** 2 <= nu])       # :373-377
        return V[:, :r], (sigma ** 2)[:r] - nu


class SVD1:
    """drivers/svd.py:126-176."""

    def __init__(self, qb):
        self.qb = qb

    def __call__(self, A, k, tol, over, rng):
        rng = np.random.default_rng(rng)
        Q, B = self.qb(A, k + over, tol, rng)
        U, s, Vh = sla.svd(B, full_matrices=False)
        if over > 0:                                            # :165-169
            c = min(k, s.size)
            U, s, Vh = U[:, :c], s[:c], Vh[:c]
        keep = ~(s < 10 * EPS)                                  # :170-174
        if not np.all(keep):
            U, s, Vh = U[:, keep], s[keep], Vh[keep]
        return Q @ U, s, Vh


class EVD1:
    """drivers/evd.py:211-289 (A symmetric)."""

    def __init__(self, qb):
        self.qb = qb

    def __call__(self, A, k, tol, over, rng):
        assert 0 < k <= min(A.shape)
        if not np.isnan(tol):
            assert 0 <= tol < np.inf
        rng = np.random.default_rng(rng)
        Q, B = self.qb(A, k + over, tol / 2, rng)               # :276
        lamb, U = sla.eigh(B @ Q)                               # :278-279
        mag = np.abs(lamb)
        r = min(k, Q.shape[1], int(np.count_nonzero(mag > 10 * EPS)))
        order = np.argsort(-mag)[:r]                            # :284
        return Q @ U[:, order], lamb[order]


# ------------------------------------------------------------------------------------------
#  Synthetic matrices used by the reference tests (parla/tests/matmakers.py:7-41)
# ------------------------------------------------------------------------------------------

def orthonormal_operator(n_rows, n_cols, rng):
    """utils/sketching.py:9-17 (sign-normalised Q of a Gaussian matrix)."""
    if n_rows < n_cols:
        return orthonormal_operator(n_cols, n_rows, rng).T
    rng = np.random.default_rng(rng)
    G = gaussian_operator(n_rows, n_cols, rng)
    Q, R = sla.qr(G, mode='economic')
    return Q * np.sign(np.diag(R))


def rand_low_rank(n_rows, n_cols, spectrum, rng, factors=False):
    """tests/matmakers.py:7-21."""
    rng = np.random.default_rng(rng)
    if isinstance(spectrum, int):
        spectrum = rng.random(size=(spectrum,))
    spectrum = np.sort(spectrum)[::-1]
    spectrum = spectrum / spectrum[0]
    rank = spectrum.size
    U = orthonormal_operator(n_rows, rank, rng)
    V = orthonormal_operator(rank, n_cols, rng)
    M = (U * spectrum) @ V
    return (M, U, spectrum, V) if factors else M


def exponent_spectrum(n_rows, n_cols, rank, rng, spectrum_param, factors=False):
    """tests/matmakers.py:34-36."""
    spec = np.exp((-np.arange(1, rank) + 1) / spectrum_param)
    return rand_low_rank(n_rows, n_cols, spec, rng, factors)


def saddle_problem(m, n, spectrum, delta, rng, rhs_scale=1.0):
    """Test problem of tests/test_drivers/test_optim/test_saddlesys.py:11-57 (make_simple_prob):
    A = U diag(spectrum) Vt, b with 70 % of its mass in range(A), Gaussian c; returns
    (A, b, c, x_opt, y_opt) with (A'A + delta I) x_opt = A'b - c and y_opt = b - A x_opt."""
    rng = np.random.default_rng(rng)
    rank = spectrum.size
    U = orthonormal_operator(m, rank, rng)
    Vt = orthonormal_operator(rank, n, rng)
    A = (U * spectrum) @ Vt
    b0 = rng.standard_normal(m)
    b_in = U @ (U.T @ b0)
    b_out = b0 - b_in
    b_in *= np.mean(spectrum) / np.linalg.norm(b_in)
    b_out *= np.mean(spectrum) / np.linalg.norm(b_out)
    b = 0.7 * b_in + (1 - 0.7) * b_out
    c = rng.standard_normal(n)
    gram = A.T @ A + delta * np.eye(n)
    rhs = A.T @ b - c
    if rhs_scale != 1.0:
        scale = rhs_scale / np.linalg.norm(rhs)
        rhs *= scale
        b *= scale
        c *= scale
    try:
        x_opt = sla.solve(gram, rhs, assume_a='pos')
    except sla.LinAlgError:
        x_opt = sla.lstsq(gram, rhs)[0]
    return A, b, c, x_opt, b - A @ x_opt


# ---------------------------------------------------------------------------------------------------------
# Test problems of the reference's own least-squares suite (restated so the GPU box, which has no
# /root/reference, can rebuild them): tests/test_drivers/test_optim/test_overdet_least_squares.py:11-112.
# Each returns (A, b, x_opt, U, s, Vt) -- the fields of the reference's AlgTestHelper (:123-139).
def simple_mat(n_rows, n_cols, scale, rng):
    """tests/matmakers.py:24-31."""
    rng = np.random.default_rng(rng)
    A = rng.normal(0, 1, (n_rows, n_cols))
    QA, RA = sla.qr(A)
    damp = 1 / np.sqrt(1 + scale * np.arange(n_cols))
    RA *= damp
    return QA @ RA


def lsq_test_problem(kind):
    """kind in {'consistent_tall', 'consistent_lowrank', 'consistent_square', 'inconsistent_orthog',
    'inconsistent_gen', 'inconsistent_stackid'}: test_overdet_least_squares.py:11-21, :23-34, :37-48, :51-68,
    :71-93, :96-112 (same seeds, same order of RNG draws)."""
    if kind == 'consistent_tall':
        rng = np.random.default_rng(190489290)
        m, n = 100, 10
        A = simple_mat(m, n, scale=1, rng=rng)
        U, s, Vt = sla.svd(A)
        x = rng.standard_normal(n)
        return A, A @ x, x, U, s, Vt
    if kind in ('consistent_lowrank', 'consistent_square'):
        low = kind == 'consistent_lowrank'
        rng = np.random.default_rng(8923890298 if low else 3278992245)
        m, n, rank = (100, 10, 5) if low else (10, 10, 10)
        U = orthonormal_operator(m, rank, rng)
        s = rng.random(rank) + 1e-4
        Vt = orthonormal_operator(rank, n, rng)
        A = (U * s) @ Vt
        x = rng.standard_normal(n)
        return A, A @ x, x, U, s, Vt
    if kind == 'inconsistent_orthog':
        n, m = 1000, 100
        rng = np.random.default_rng(19837647834763)
        U = orthonormal_operator(n, m, rng)
        Vt = orthonormal_operator(m, m, rng)
        s = rng.random(m) + 1e-4
        A = (U * s) @ Vt
        b = rng.standard_normal(n)
        b = b - U @ (U.T @ b)
        b *= 1e2 / sla.norm(b)
        return A, b, np.zeros(m), U, s, Vt
    if kind == 'inconsistent_gen':
        rng = np.random.default_rng(897809809)
        m, n, num_hi = 1000, 100, 30
        num_lo = n - num_hi
        hi_spec = 1e5 * np.ones(num_hi) + rng.random(num_hi)
        lo_spec = np.ones(num_lo) + rng.random(num_lo)
        spec = np.concatenate([hi_spec, lo_spec])
        U = orthonormal_operator(m, n, rng)
        Vt = orthonormal_operator(n, n, rng)
        A = (U * spec) @ Vt
        hi_x = rng.standard_normal(num_hi) / 1e5
        lo_x = rng.standard_normal(num_lo)
        x = np.concatenate([hi_x, lo_x])
        b_orth = rng.standard_normal(m) * 1e2
        b_orth -= U @ (U.T @ b_orth)
        return A, A @ x + b_orth, x, U, spec, Vt
    if kind == 'inconsistent_stackid':
        rng = np.random.default_rng(2837592038243)
        A = np.tile(np.eye(70), (10, 1))
        m, n = A.shape
        A = (A.T * rng.lognormal(size=(m,))).T
        x0 = rng.standard_normal(size=(n,))
        b = A @ x0 + rng.standard_normal(size=(m,))
        U, spec, Vt = sla.svd(A, full_matrices=False)
        x = Vt.T @ (U.T @ b) / spec
        return A, b, x, U, spec, Vt
    raise ValueError(kind)


def consistent_lowrank_problem():
    A, b, x, _, _, Vt = lsq_test_problem('consistent_lowrank')
    return A, b, x, Vt


def loglinear_fit(x, y):
    """utils/stats.py:6-29: least-squares fit of log(y) ~ a + b x; returns ([a, b], R^2)."""
    x = np.asarray(x, float).ravel()
    y = np.asarray(y, float).ravel()
    keep = y > 0
    x, logy = x[keep], np.log(y[keep])
    mat = np.column_stack([np.ones(x.size), x])
    fit = sla.lstsq(mat, logy)[0]
    ss_tot = np.sum((logy - np.mean(logy)) ** 2)
    ss_res = np.sum((logy - mat @ fit) ** 2)
    return fit, 1 - ss_res / ss_tot
