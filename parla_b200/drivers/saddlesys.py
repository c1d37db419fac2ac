"""Sketch-and-precondition drivers for saddle-point systems on B200.

Mirrors parla/drivers/saddlesys.py: ``SaddleSolver`` (:20-73), ``sps`` (:76-86), ``SPS1`` (:89-223,
SVD / Nystrom preconditioner + PCG) and ``SPS2`` (:226-327, reduction to over-determined least squares
+ LSQR).  Same constructors and ``__call__(A, b, c, delta, tol, iter_lim, rng, logging)`` signatures;
``A`` / ``b`` are torch CUDA fp64 tensors (or ``parallel.RowSharded`` row shards), ``c`` is a replicated
device n-vector; numpy inputs are uploaded and the results come back as numpy.

    [  I   |     A   ] [y] = [b]            (A' A + delta I) x = A' b - c
    [  A'  | -delta*I] [x]   [c]

What differs from the reference's execution, not from its mathematics:
  * ``[A; sqrt(delta) I]`` is never materialised (the reference copies A, saddlesys.py:286); the ridge
    rows live in n-vectors (comps/preconditioning.py);
  * the SVD of the (lifted) sketch goes through its Householder QR -- only an n x n SVD is left to cuSOLVER;
  * the Gram product A'(A p) of PCG and the presolve tests are single fused passes over A.
"""
import math
import time

import numpy as np
import torch

from .. import distla
from .. import kernels as K
from ..comps.determiter.logging import SketchAndPrecondLog
from ..comps.determiter import saddle as dsad
from ..comps import preconditioning as rpc
from ..comps.sketchers import oblivious as sko
from ..parallel import RowSharded, allreduce_, unwrap
from ..utils.sketching import as_device_operator
from .least_squares import _clock, _sketch, dim_checks

F64 = torch.float64


class SaddleSolver:
    """Interface of saddlesys.py:20-73."""

    def __call__(self, A, b, c, delta, tol, iter_lim, rng, logging):
        raise NotImplementedError()

    exec = __call__


def sps(A, b, c, delta, tol, iter_lim, rng, sampling_factor=3, vec_nnz=8, method='pcg'):
    """saddlesys.py:76-86."""
    skop = sko.SkOpSJ(vec_nnz)
    if method == 'pcg':
        solver = dsad.PcSS1()
    elif method == 'lsqr':
        solver = dsad.PcSS2()
    else:
        raise ValueError(f'Method {method} not recognized. Use "pcg" or "lsqr".')
    alg = SPS1(skop, sampling_factor, solver)
    return alg(A, b, c, delta, tol, iter_lim, rng, logging=True)


def _upload(A, b, c):
    """-> (A, b, c, host): numpy inputs are uploaded to the current device."""
    host = isinstance(A, np.ndarray)
    if host:
        dev = torch.device("cuda", torch.cuda.current_device())
        up = lambda v: None if v is None else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
        A, b, c = up(A), up(b), up(c)
    return A, b, c, host


def _download(host, *vals):
    if not host:
        return vals
    return tuple(v.cpu().numpy() if isinstance(v, torch.Tensor) else v for v in vals)


def _mv(M, z):
    """M @ z for a small dense row-major M (streaming kernel; M is L2 resident)."""
    out = torch.zeros(M.shape[0], dtype=F64, device=M.device)
    K.stream_pass(M, w=z, u=out, sa=1.0, su=0.0, flags=K.PASS_DOT)
    return out


def _mtv(M, w):
    """M^T @ w."""
    return K.rmatvec(M, w)[:M.shape[1]].clone()


def _norm(v):
    return math.sqrt(float(K.sumsq(v)))


def _svd_of_tall(Y):
    """(sigma, Wt) of the thin SVD of a tall (possibly row-sharded) matrix, through R of its QR."""
    if isinstance(Y, RowSharded):
        import torch.distributed as dist
        group = Y.group if Y.group is not None else dist.group.WORLD
        k = Y.local.shape[1]
        _, R1 = K.qr_economic(Y.local)
        stack = torch.empty(dist.get_world_size(group) * k, k, dtype=F64, device=R1.device)
        dist.all_gather_into_tensor(stack, R1.contiguous(), group=group)
        _, R = K.qr_economic(stack)
    else:
        W = Y.clone()
        k = W.shape[1]
        K.geqrf(W, k)
        R = torch.triu(W[:k, :k])
    _, sigma, Wt = torch.linalg.svd(R, full_matrices=False, driver='gesvd')
    return sigma, Wt


class SPS1(SaddleSolver):
    """SVD-based sketch-and-precondition for saddle systems (saddlesys.py:89-223).  With
    ``sampling_factor < 1`` the preconditioner is a rank-d Nystrom-type approximation
    (``nystrom_strategy`` 'left' or 'right', :158-184)."""

    NYSTROM_STRATEGIES = {'left', 'right'}

    def __init__(self, sketch_op_gen, sampling_factor, iterative_solver=None):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        if iterative_solver is None:
            iterative_solver = dsad.PcSS1()
        self.iterative_solver = iterative_solver
        self.nystrom_strategy = 'left'

    def __call__(self, A, b, c, delta, tol, iter_lim, rng, logging=True):
        A, b, c, host = _upload(A, b, c)
        m, n = A.shape
        d = int(self.sampling_factor * n)                                     # :134
        rng = np.random.default_rng(rng)
        assert self.nystrom_strategy in self.NYSTROM_STRATEGIES
        A_loc, _, group = unwrap(A)
        dev = A_loc.device
        if b is None:
            b = torch.zeros(A_loc.shape[0], dtype=F64, device=dev)
            if isinstance(A, RowSharded):
                b = RowSharded(b, A.row_offset, A.m_global, A.group)
        b_loc = unwrap(b)[0]

        quick_time = _clock(logging)
        log = SketchAndPrecondLog()
        nystrom_like = d < n
        sigma = Vh = None
        if not nystrom_like:                                                  # :146-157
            tic = quick_time()
            _, W = _sketch(self.sketch_op_gen, d, A, None, delta, rng)
            log.time_sketch = quick_time() - tic
            tic = quick_time()
            K.geqrf(W, n)                       # svd(A_ske) = (Q U_r, sigma, Vh) with R_qr = U_r sigma Vh
            M, _U_r, sigma, Vh = rpc.svd_right_precond(torch.triu(W[:n, :n]))
            log.time_factor = quick_time() - tic
        elif self.nystrom_strategy == 'right':                                # :159-170
            tic = quick_time()
            S = as_device_operator(self.sketch_op_gen(n, d, rng), dev).to_dense()
            A_sample = distla.mm(A, S)
            log.time_sketch = quick_time() - tic
            tic = quick_time()
            Q = distla.orth(A_sample)
            A_ske = distla.mm_t(Q, A)
            _, sig, Vh_ = torch.linalg.svd(A_ske, full_matrices=False, driver='gesvd')
            M = (Vh_.T / sig).contiguous()
            log.time_factor = quick_time() - tic
        else:                                                                 # :171-184
            tic = quick_time()
            _, W = _sketch(self.sketch_op_gen, d, A, None, 0.0, rng)
            log.time_sketch = quick_time() - tic
            tic = quick_time()
            V = K.qr_economic(W[:d, :n].T.contiguous())[0]                    # orth(A_ske.T), n x d
            sig, Wt = _svd_of_tall(distla.mm(A, V))
            M = K.gemm(V, (Wt.T / sig).contiguous())
            log.time_factor = quick_time() - tic

        gram = dsad.GramOperator(A)
        rhs = gram.rmatvec(b_loc)                                             # :187-189
        if c is not None:
            rhs = rhs - c
        # Presolve (:196-208).  z_ske = Sigma^+ V' rhs is the same vector as rhs_pc = M' rhs.
        tic = quick_time()
        z_ske = None
        if not nystrom_like:
            z_ske = _mv(Vh, rhs) / sigma
            x_ske = _mv(M, z_ske)
            lhs = gram(x_ske)[:n] + delta * x_ske
            gap = _mtv(M, lhs - rhs)
            if _norm(gap) >= _norm(z_ske):
                z_ske = None
        log.time_presolve = quick_time() - tic

        tic = quick_time()
        if isinstance(self.iterative_solver, dsad.PcSS1):
            res = self.iterative_solver(A, b, c, delta, tol, iter_lim, M, False, z_ske, _rhs=rhs)
        else:
            res = self.iterative_solver(A, b, c, delta, tol, iter_lim, M, False, z_ske)
        log.time_iterate = quick_time() - tic
        x_star, y_star = res[0], res[1]

        if logging:                                                           # :218-220
            log.wrap_up(res[2], _norm(rhs))
            log.error_desc = self.iterative_solver.ERROR_METRIC_INFO
            log.iters = int(np.atleast_1d(res[2]).size)
        x_star, y_star = _download(host, x_star, y_star)
        return x_star, y_star, log

    exec = __call__


class SPS2(SaddleSolver):
    """Sketch, reduce to over-determined least squares, precondition by the SVD of the sketch, LSQR
    (saddlesys.py:226-327)."""

    def __init__(self, sketch_op_gen, sampling_factor, iterative_solver=None):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        if iterative_solver is None:
            iterative_solver = dsad.PcSS2()
        self.iterative_solver = iterative_solver

    def __call__(self, A, b, c, delta, tol, iter_lim, rng, logging=False):
        A, b, c, host = _upload(A, b, c)
        m, n = A.shape
        sqrt_delta = math.sqrt(delta)
        d = dim_checks(self.sampling_factor, m, n)
        rng = np.random.default_rng(rng)
        A_loc, row_off, group = unwrap(A)
        dev = A_loc.device
        m_loc = A_loc.shape[0]
        b_loc = torch.zeros(m_loc, dtype=F64, device=dev) if b is None else unwrap(b)[0]

        quick_time = _clock(logging)
        log = SketchAndPrecondLog()

        # Sketch the data matrix (+ ridge rows)                                 :275-279
        tic = quick_time()
        S, W = _sketch(self.sketch_op_gen, d, A, None, delta, rng)
        log.time_sketch = quick_time() - tic

        # Factor the sketch: thin SVD of the lifted sketch through its QR       :282-284
        tic = quick_time()
        Q, R_qr = K.qr_economic(W[:, :n])
        M, U_r, sigma, Vh = rpc.svd_right_precond(R_qr)
        U = K.gemm(Q, U_r.contiguous())                                        # (d [+ n]) x rank
        log.time_factor = quick_time() - tic

        # Convert to over-determined least squares                              :287-295
        tic = quick_time()
        b_top = b_loc.clone()
        b_ridge = torch.zeros(n, dtype=F64, device=dev) if delta > 0 else None
        if c is not None and _norm(c) > 0:
            # v = pinv(A_ske_aug') c must satisfy A_ske_aug' v = c to working accuracy (it defines the transformed
            # right-hand side).  With full column rank that is v = Q R^{-T} c -- one triangular solve against
            # the Householder R, exact to eps * cond, whichever route produced (M, U, sigma, Vh); U from the
            # Gram/eigh route is orthonormal only to eps * cond^2 and is used for the presolve alone.
            if M.shape[1] == n:
                v = _mv(Q, K.trsv_upper(R_qr, c, trans=True))
            else:
                v = _mv(U, _mv(Vh, c) / sigma)
            b_top -= S.rmatvec(v[:d].contiguous(), m_local=m_loc, row_offset=row_off)
            if delta > 0:
                b_ridge -= v[d:]
        log.time_convert = quick_time() - tic

        # Presolve: z_ske = U' [S b_top; b_ridge], accepted if it beats the zero vector   :298-304
        tic = quick_time()
        sb = torch.zeros(U.shape[0], dtype=F64, device=dev)
        sb_top = torch.zeros(d, 2, dtype=F64, device=dev)
        S.sketch_into(b_top.reshape(-1, 1), None, sb_top[:, :1], row_offset=row_off)
        sb[:d] = allreduce_(sb_top, group)[:, 0]
        if delta > 0:
            sb[d:] = b_ridge
        z_ske = _mtv(U, sb)
        op = rpc.PrecondOperator(A, delta, M, False)
        bsq = float(allreduce_(K.sumsq(b_top), group)) + (float(K.sumsq(b_ridge)) if delta > 0 else 0.0)
        warm = dict(u=b_top.clone(), ub=b_ridge.clone() if delta > 0 else None,
                    zss=torch.empty(n + 1, dtype=F64, device=dev),
                    t=torch.empty(M.shape[1], dtype=F64, device=dev))
        xw = torch.empty(n, dtype=F64, device=dev)
        op.bidiag_pass(z_ske, warm["u"], warm["ub"], warm["zss"], xw, warm["t"], sa=-1.0, su=1.0)
        if math.sqrt(float(warm["zss"][n])) >= math.sqrt(bsq):
            z_ske, warm = None, None
        log.time_presolve = quick_time() - tic

        # Main iterative phase                                                  :307-313
        tic = quick_time()
        if isinstance(self.iterative_solver, dsad.PcSS2):
            res = self.iterative_solver(A, b_top, None, delta, tol, iter_lim, M, False, z_ske, _op=op, _warm=warm,
                                        _need_y=False, _b_ridge=b_ridge)
        else:
            raise NotImplementedError("SPS2 on the device needs the LSQR-backed PcSS2 (the ridge rows of the "
                                      "transformed system are kept implicit)")
        log.time_iterate = quick_time() - tic
        x_star = res[0]
        y_star = op.residual_and_atb(x_star, b_loc)                            # b - A x with the ORIGINAL b (:311)

        if logging:                                                            # :316-321
            g = op.rmatvec_plain(b_top)
            if delta > 0:
                g = g + sqrt_delta * b_ridge
            log.wrap_up(res[2], _norm(op.precond_t(g)))
            log.error_desc = self.iterative_solver.ERROR_METRIC_INFO
            log.error_desc += "The metric above is computed w.r.t. a transformed problem."
            log.iters = int(np.atleast_1d(res[2]).size)
            log.passes_over_A = op.passes + 1
        x_star, y_star = _download(host, x_star, y_star)
        return x_star, y_star, log

    exec = __call__
