"""Sketch-and-precondition / sketch-and-solve drivers for over-determined least squares on B200.

Mirrors parla/drivers/least_squares.py: ``OverLstsqSolver`` (:16-86), ``dim_checks`` (:89-104),
``SSO1`` (:114-189) and ``SPO`` (:206-369; "SAP1" == mode 'qr', "SAP2" == mode 'svd').  Same
constructor and ``__call__(A, b, delta, tol, iter_lim, rng, logging=True)`` signatures; ``A`` / ``b``
are torch CUDA fp64 tensors (or ``parallel.RowSharded`` shards, or numpy arrays which are uploaded
and whose result is returned as numpy).  ``exec`` is an alias of ``__call__`` (RandLAPACK naming).

What differs from the reference's execution, not from its mathematics:
  * ``S @ A`` and ``S @ b`` are one kernel writing ``[A_ske | b_ske]`` into one d x (n+1) buffer; the
    Householder QR of its first n columns then leaves ``Q^T b_ske`` in the last column
    (least_squares.py:311-315 in one factorisation, Q never formed);
  * the presolve residual ``A x_ske - b`` (:344) IS LSQR's initial residual (lsqr.py:367-370), so it is
    computed once and handed to the iterative solver;
  * ``y = b - A x`` (saddle.py:199) and ``A^T b`` for the log (:361) share one pass over A.
"""
import math
import os
import time
import warnings

import numpy as np
import torch

from .. import kernels as K
from ..comps.determiter.logging import SketchAndPrecondLog
from ..comps.determiter import saddle as dsad
from ..comps import preconditioning as rpc
from ..parallel import RowSharded, allreduce_, unwrap
from ..utils.sketching import as_device_operator, shard_context

F64 = torch.float64
EXPLICIT_INVERSE = os.environ.get("PLA_EXPLICIT_RINV", "1") != "0"


class OverLstsqSolver:
    """min ||A x - b||_2^2 + delta ||x||_2^2 for a tall A.  Interface of least_squares.py:16-86."""

    def __call__(self, A, b, delta, tol, iter_lim, rng):
        raise NotImplementedError()

    exec = __call__


def dim_checks(sampling_factor, n_rows, n_cols):
    """least_squares.py:89-104."""
    assert n_rows >= n_cols
    d = int(sampling_factor * n_cols)
    if d > n_rows:
        msg = f"""
        The embedding dimension "d" should not be larger than the
        number of rows of the data matrix. Here, an embedding dimension
        of d={d} has been requested for a matrix with only {n_rows} rows.
        We will proceed by setting d={n_rows}. This parameter choice will
        result in a very inefficient algorithm!
        """
        warnings.warn(msg)
        d = n_rows
    assert d >= n_cols
    return d


def _to_device(A, b, defer=False):
    """Accept HOST buffers (numpy arrays, CPU torch tensors or RowSharded shards of those; ideally pinned).
    -> (A, b, host_kind).  With ``defer`` the upload is left to :func:`_sketch`, which streams it in row blocks on a
    copy stream while the blocks already on the device are being sketched."""
    host = None
    A_loc, b_loc = unwrap(A)[0], (None if b is None else unwrap(b)[0])
    if isinstance(A_loc, np.ndarray):
        host = "numpy"
        A_loc = torch.from_numpy(np.ascontiguousarray(A_loc, dtype=np.float64))
        b_loc = None if b_loc is None else torch.from_numpy(np.ascontiguousarray(b_loc, dtype=np.float64))
    elif isinstance(A_loc, torch.Tensor) and not A_loc.is_cuda:
        host = "torch"
    if host is None:
        return A, b, None
    if not defer:
        dev = torch.device("cuda", torch.cuda.current_device())
        A_loc = A_loc.to(dev, non_blocking=True)
        b_loc = None if b_loc is None else b_loc.to(dev, non_blocking=True)
    return _like_input(A, A_loc), (None if b is None else _like_input(b, b_loc)), host


def _like_input(orig, local):
    """Re-wrap a local block the way ``orig`` was given (RowSharded or plain)."""
    if isinstance(orig, RowSharded):
        return RowSharded(local, orig.row_offset, orig.m_global, orig.group)
    return local


def _to_host(x, host):
    if host == "numpy":
        return x.cpu().numpy()
    if host == "torch":
        return x.cpu()
    return x


UPLOAD_BLOCKS = int(os.environ.get("PLA_UPLOAD_BLOCKS", "8"))
_COPY_STREAMS = {}


def _streamed_sketch(S, A_host, b_host, W, row_off, stats):
    """Upload a host-resident block of rows in ``UPLOAD_BLOCKS`` pieces on a copy stream and sketch every piece as
    soon as it has arrived (``W += S[:, piece] [A | b][piece]``): the sketch costs no time on top of the PCIe
    transfer.  Returns the device copies (A, b); ``stats`` receives bytes and the copy stream's seconds."""
    dev = W.device
    m_loc, n = A_host.shape
    main = torch.cuda.current_stream()
    key = torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream()
    cs = _COPY_STREAMS[key]
    A_dev = torch.empty(m_loc, n, dtype=F64, device=dev)
    b_dev = None if b_host is None else torch.empty(m_loc, dtype=F64, device=dev)
    step = max(4096, -(-m_loc // max(1, UPLOAD_BLOCKS)) // 4096 * 4096)       # multiple of 4096 rows
    bounds = [(r0, min(r0 + step, m_loc)) for r0 in range(0, m_loc, step)]
    cs.wait_stream(main)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    arrived = []
    with torch.cuda.stream(cs):
        t0.record()
        for r0, r1 in bounds:
            A_dev[r0:r1].copy_(A_host[r0:r1], non_blocking=True)
            if b_dev is not None:
                b_dev[r0:r1].copy_(b_host[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            arrived.append(ev)
        t1.record()
    for i, (r0, r1) in enumerate(bounds):
        main.wait_event(arrived[i])
        piece = S.column_slice(r0, r1 - r0)
        piece.sketch_into(A_dev[r0:r1], None if b_dev is None else b_dev[r0:r1], W, row_offset=row_off + r0,
                          accumulate=i > 0)
    stats["bytes"] = m_loc * n * 8 + (0 if b_host is None else m_loc * 8)
    stats["events"] = (t0, t1)
    return A_dev, b_dev


def _sketch(sketch_op_gen, d, A, b, delta, rng, upload=None):
    """[A_ske | b_ske] (+ ridge rows) in one (d [+ n]) x (n + 1) row-major buffer.  With ``upload`` (a dict) a
    host-resident A / b is uploaded here, overlapped with the sketch; the device copies are returned in it."""
    A_loc, row_off, group = unwrap(A)
    b_loc = None if b is None else unwrap(b)[0]
    m, n = A.shape
    on_host = not A_loc.is_cuda
    dev = torch.device("cuda", torch.cuda.current_device()) if on_host else A_loc.device
    with shard_context(row_off, A_loc.shape[0]):
        S = as_device_operator(sketch_op_gen(d, m, rng), dev)
    if S.shape[1] != A_loc.shape[0]:           # a full-width operator (e.g. a replayed reference S)
        S = S.column_slice(row_off, A_loc.shape[0])
    d_aug = d + (n if delta > 0 else 0)
    # even leading dimension (16-byte aligned rows) so the QR's tensor-core trailing updates and
    # the sketch kernels can use vector accesses; the pad column is kept at zero
    ld = n + 1 + ((n + 1) % 2)
    Wfull = torch.zeros(d_aug, ld, dtype=F64, device=dev)
    W = Wfull[:, :n + 1]
    if on_host:
        from ..utils.sketching import GaussianOperator, SJLTOperator
        if isinstance(S, (GaussianOperator, SJLTOperator)) and A_loc.shape[0] >= 2 * 4096:
            A_dev, b_dev = _streamed_sketch(S, A_loc, b_loc, W[:d], row_off, upload)
        else:                                   # other operators need the whole block: plain upload, then sketch
            A_dev = A_loc.to(dev, non_blocking=True)
            b_dev = None if b_loc is None else b_loc.to(dev, non_blocking=True)
            S.sketch_into(A_dev, b_dev, W[:d], row_offset=row_off)
        upload["A"], upload["b"] = _like_input(A, A_dev), (None if b is None else _like_input(b, b_dev))
    else:
        S.sketch_into(A_loc, b_loc, W[:d], row_offset=row_off)
    allreduce_(Wfull[:d], group)
    if delta > 0:
        W[d:, :].zero_()
        W[d:, :n].diagonal().fill_(math.sqrt(delta))
    return S, W


def _factor_sketch(W, n, group):
    """Householder QR of the (replicated) sketch buffer: column-distributed over the ranks when A is row-sharded and
    the sketch is large enough, the single-GPU factorisation otherwise."""
    from .. import distla
    if os.environ.get("PLA_QR_DIST", "1") != "0" and distla.geqrf_distributed_ok(W.shape[0], n, group):
        return distla.geqrf_distributed(W, n, group)
    return K.geqrf(W, n)


def sso1(A, b, delta, rng, sampling_factor=3, vec_nnz=8, lapack_driver='gelsd'):
    """least_squares.py:108-111."""
    from ..comps.sketchers import oblivious as sko
    alg = SSO1(sko.SkOpSJ(vec_nnz), sampling_factor, lapack_driver, overwrite_sketch=True)
    return alg(A, b, delta, np.nan, 1, rng, logging=True)


class SSO1(OverLstsqSolver):
    """Sketch-and-solve (least_squares.py:114-189): x = argmin ||S A x - S b||, via Householder QR."""

    def __init__(self, sketch_op_gen, sampling_factor, lapack_driver=None, overwrite_sketch=True):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.lapack_driver = lapack_driver
        self.overwrite_sketch = overwrite_sketch

    def __call__(self, A, b, delta, tol, iter_lim, rng, logging=True):
        if not np.isnan(tol):
            warnings.warn("""
            This OverLstsqSolver implementation cannot directly control
            approximation error. Parameter "tol" is being ignored.
            """)
        if iter_lim > 1:
            warnings.warn("""
            This OverLstsqSolver implementation is not iterative.
            Parameter "iter_lim" is being ignored.
            """)
        A, b, host = _to_device(A, b)
        n_rows, n_cols = A.shape
        d = dim_checks(self.sampling_factor, n_rows, n_cols)
        rng = np.random.default_rng(rng)
        log = {'time_sketch': -1.0, 'time_solve': -1.0}
        quick_time = _clock(logging)
        tic = quick_time()
        _, W = _sketch(self.sketch_op_gen, d, A, b, delta, rng)
        log['time_sketch'] = quick_time() - tic
        tic = quick_time()
        # la.lstsq(A_ske, b_ske, lapack_driver=...) (:184-186).  Householder QR + one triangular solve when the
        # sketch has full numerical rank; otherwise the minimum-norm solution through the SVD of R with
        # gelsd's cut-off (every scipy driver -- gelsd, gelsy, gelss -- is rank revealing, so all map here).
        K.geqrf(W, n_cols)
        R = W[:n_cols, :n_cols]
        dabs = R.diagonal().abs()
        dmin, dmax = (float(v) for v in torch.stack((dabs.min(), dabs.max())).cpu())
        if dmin > torch.finfo(F64).eps * n_cols * dmax:
            x_ske = K.trsv_upper(R, W[:n_cols, n_cols].contiguous())
        else:
            M, U_r, _, _ = rpc.svd_right_precond(torch.triu(R), exact=True)
            x_ske = M @ (U_r.T @ W[:n_cols, n_cols])
        log['time_solve'] = quick_time() - tic
        return _to_host(x_ske, host), log

    exec = __call__


def _clock(logging):
    if not logging:
        return lambda: 0
    def now():
        torch.cuda.synchronize()
        return time.time()
    return now


def spo1(A, b, delta, tol, iter_lim, rng, sampling_factor=3, vec_nnz=8):
    """least_squares.py:193-196 (SVD preconditioner)."""
    from ..comps.sketchers import oblivious as sko
    return SPO(sko.SkOpSJ(vec_nnz), sampling_factor, mode='svd')(A, b, delta, tol, iter_lim, rng, logging=True)


def spo3(A, b, delta, tol, iter_lim, rng, sampling_factor=3, vec_nnz=8, mode='qr'):
    """least_squares.py:200-203."""
    from ..comps.sketchers import oblivious as sko
    return SPO(sko.SkOpSJ(vec_nnz), sampling_factor, mode)(A, b, delta, tol, iter_lim, rng, logging=True)


class SPO(OverLstsqSolver):
    """Sketch-and-precondition with LSQR (least_squares.py:206-369)."""

    def __init__(self, sketch_op_gen, sampling_factor: int, mode='qr'):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.mode = mode
        self.iterative_solver = dsad.PcSS2()  # implements LSQR

    def __call__(self, A, b, delta, tol, iter_lim, rng, logging=True):
        A, b, host = _to_device(A, b, defer=True)
        n_rows, n_cols = A.shape
        sqrt_delta = math.sqrt(delta)
        d = dim_checks(self.sampling_factor, n_rows, n_cols)
        rng = np.random.default_rng(rng)

        quick_time = _clock(logging)
        log = SketchAndPrecondLog()

        # Sketch the data matrix (and the right-hand side, same kernel)            :302-303, :314
        # (host-resident A / b: uploaded here in row blocks, each block sketched while the next one is in flight)
        tic = quick_time()
        upload = {} if host is not None else None
        S, W = _sketch(self.sketch_op_gen, d, A, b, delta, rng, upload=upload)
        if host is not None:
            A, b = upload.pop("A"), upload.pop("b")
            self._upload_stats = upload          # (bytes + the copy stream's events only: no reference to A)
        log.time_sketch = quick_time() - tic
        A_loc, _, group = unwrap(A)
        b_loc = unwrap(b)[0]
        dev = A_loc.device

        # Factor the sketch; sketch-and-solve presolve                              :306-341
        n = n_cols
        if self.mode == 'qr':
            tic = quick_time()
            _factor_sketch(W, n, group)                     # R = triu(W[:n,:n]), W[:, n] = Q^T [b_ske; 0]
            R = W[:n, :n]
            log.time_factor = quick_time() - tic
            tic = quick_time()
            z_ske = W[:n, n].contiguous()
            tri = True
        elif self.mode == 'chol':
            tic = quick_time()
            A_ske = W[:d, :n]
            G = K.gemm(A_ske, A_ske, transa=True)
            if delta > 0:
                G.diagonal().add_(delta)
            R = torch.linalg.cholesky(G, upper=True)        # small n x n factorisation: cuSOLVER glue
            log.time_factor = quick_time() - tic
            tic = quick_time()
            g = K.rmatvec(A_ske, W[:d, n].contiguous())[:n]
            z_ske = K.trsv_upper(R, g, trans=True)
            tri = True
        elif self.mode == 'svd':
            # SVD of the sketch through its Householder QR (same factorisation as mode 'qr'): A_ske = Q R_qr,
            # R_qr = U_r diag(sigma) Vh  =>  svd(A_ske) = (Q U_r, sigma, Vh).  Only the n x n SVD goes to
            # cuSOLVER, and U^T b_ske = U_r^T (Q^T b_ske) comes out of the same QR (last column of W).
            tic = quick_time()
            _factor_sketch(W, n, group)
            X = None
            if n >= rpc.FAST_SVD_MIN_N and rpc.FAST_SVD:
                # full numerical rank with a wide margin: R_qr^{-1} = (V / sigma) U_r^T is the SVD preconditioner up
                # to an orthogonal factor on the right -- same x, same error history, no n x n decomposition
                X = rpc.inverse_if_well_conditioned(W[:n, :n])
            if X is not None:
                R = X
                log.time_factor = quick_time() - tic
                tic = quick_time()
                z_ske = W[:n, n].contiguous()
            else:
                R, U_r, sigma, Vh = rpc.svd_right_precond(torch.triu(W[:n, :n]))   # :330-339
                log.time_factor = quick_time() - tic
                tic = quick_time()
                z_ske = K.rmatvec(U_r, W[:n, n].contiguous())[:R.shape[1]].clone()  # U[:d].T @ b_ske
            tri = False
        else:
            raise ValueError()

        # The two triangular solves per LSQR iteration (preconditioning.py:28,37) are sequential and
        # latency-bound; with the explicit inverse they become two bandwidth-bound matvecs over an
        # L2-resident matrix.  Any fixed nonsingular M gives the same minimiser x = M z, so this only
        # perturbs the preconditioner by O(cond(R) eps).  (z_ske above is still solved against R.)
        M_pc, tri_pc = R, tri
        if tri and n <= K.PASS_MAX_N and EXPLICIT_INVERSE:
            M_pc, tri_pc = K.trtri_upper(R), False
            log.time_factor += quick_time() - tic
            tic = quick_time()

        # Presolve acceptance test (:344-352).  b - A x_ske is LSQR's starting residual, so the pass
        # doubles as the solver's initialisation.
        op = rpc.PrecondOperator(A, delta, M_pc, tri_pc)
        bnorm = math.sqrt(float(allreduce_(K.sumsq(b_loc), group)))
        warm = dict(u=b_loc.clone(),
                    ub=torch.zeros(n, dtype=F64, device=dev) if delta > 0 else None,
                    zss=torch.empty(n + 1, dtype=F64, device=dev),
                    t=torch.empty(R.shape[1], dtype=F64, device=dev))
        xw = torch.empty(n, dtype=F64, device=dev)
        op.bidiag_pass(z_ske, warm["u"], warm["ub"], warm["zss"], xw, warm["t"], sa=-1.0, su=1.0)
        with np.errstate(divide='ignore', invalid='ignore'):       # b == 0 gives nan, as in the reference (:347)
            rel_err = float(np.float64(math.sqrt(float(warm["zss"][n]))) / np.float64(bnorm))
        if rel_err >= 1 or (rel_err > 1e-15 and R.shape[0] != R.shape[1]):
            # Either the zero vector is a better solution, or we have an inconsistent
            # rank-deficient problem (which forces us to initialize at the origin).
            z_ske, warm = None, None
        log.time_presolve = quick_time() - tic

        # Iterative phase                                                           :355-358
        tic = quick_time()
        # y = b - A x is not part of SPO's result (least_squares.py:369 returns res[0] only); the pass
        # that produces it is kept only when logging needs A^T b from the same read of A.
        need_pass = bool(logging) and not (z_ske is None)
        res = self.iterative_solver(A, b, None, delta, tol, iter_lim, M_pc, tri_pc, z_ske, _op=op, _warm=warm,
                                    _need_y=need_pass)
        log.time_iterate = quick_time() - tic

        if logging:                                                               # :360-367
            ar0 = op.precond_t(op.atb)
            iter_errors = res[2]
            log.wrap_up(iter_errors, float(torch.linalg.vector_norm(ar0)))
            log.error_desc = self.iterative_solver.ERROR_METRIC_INFO
            log.iters = int(np.atleast_1d(iter_errors).size)
            log.passes_over_A = op.passes + 1      # + the sketch
        self.last_residual = res[1]
        self.iterative_solver.last_op = None     # (it references A: do not keep a 64 GiB upload alive after the call)
        x = res[0]
        return _to_host(x, host), log

    exec = __call__

    @property
    def last_upload(self):
        """{'bytes', 'seconds'} of the last host -> device upload done inside a call (None for device inputs)."""
        st = getattr(self, "_upload_stats", None)
        if not st or "events" not in st:
            return None
        t0, t1 = st["events"]
        t1.synchronize()
        return {"bytes": st["bytes"], "seconds": t0.elapsed_time(t1) * 1e-3}


class SAP1(SPO):
    """"SAP1" of the reference's change log (CHANGELOG.md:57): sketch-and-precondition with the QR
    preconditioner, i.e. ``SPO(sketch_op_gen, sampling_factor, mode='qr')`` (least_squares.py:306-316)."""

    def __init__(self, sketch_op_gen, sampling_factor):
        super().__init__(sketch_op_gen, sampling_factor, mode='qr')


class SAP2(SPO):
    """"SAP2" (CHANGELOG.md:57): the SVD preconditioner, ``SPO(..., mode='svd')`` (least_squares.py:330-339);
    handles rank-deficient sketches."""

    def __init__(self, sketch_op_gen, sampling_factor):
        super().__init__(sketch_op_gen, sampling_factor, mode='svd')


class UnderLstsqSolver:
    """min ||y|| s.t. A' y = c for a tall A.  Interface of least_squares.py:372-417."""

    def __call__(self, A, c, tol, iter_lim, rng, logging=False):
        raise NotImplementedError()

    exec = __call__


def spu1(A, c, tol, iter_lim, rng, sampling_factor=3, vec_nnz=8):
    """least_squares.py:419-422."""
    from ..comps.sketchers import oblivious as sko
    return SPU1(sko.SkOpSJ(vec_nnz), sampling_factor)(A, c, tol, iter_lim, rng, logging=True)


class SPU1(UnderLstsqSolver):
    """SVD-based sketch-and-precondition for under-determined least squares
    (least_squares.py:425-494): sketch A, M = V / sigma from the SVD of the sketch, then LSQR on
    (A M)^T (PcSS2's under-determined branch)."""

    def __init__(self, sketch_op_gen, sampling_factor: int):
        self.sketch_op_gen = sketch_op_gen
        self.sampling_factor = sampling_factor
        self.iterative_solver = dsad.PcSS2()  # implements LSQR

    def __call__(self, A, c, tol, iter_lim, rng, logging=True):
        host = None
        if isinstance(A, np.ndarray):
            host = "numpy"
            dev = torch.device("cuda", torch.cuda.current_device())
            A = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float64)).to(dev)
            c = torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64)).to(dev)
        n_rows, n_cols = A.shape
        d = dim_checks(self.sampling_factor, n_rows, n_cols)
        rng = np.random.default_rng(rng)
        A_loc, _, group = unwrap(A)

        quick_time = _clock(logging)
        log = SketchAndPrecondLog()

        tic = quick_time()                                                    # :467-471
        _, W = _sketch(self.sketch_op_gen, d, A, None, 0.0, rng)
        log.time_sketch = quick_time() - tic

        tic = quick_time()                                                    # :474-476
        M, U, sigma, Vh = rpc.svd_right_precond(W[:, :n_cols])
        log.time_factor = quick_time() - tic

        tic = quick_time()                                                    # :479-483
        res = self.iterative_solver(A, None, c, 0.0, tol, iter_lim, M, False, None)
        log.time_iterate = quick_time() - tic
        y_star = res[1]

        if logging:                                                           # :485-492
            op = self.iterative_solver.last_op
            Av = op.matvec_plain(op.precond(op.precond_t(c)))
            nrm = math.sqrt(float(allreduce_(K.sumsq(Av), group)))
            log.wrap_up(res[2], nrm)
            log.iters = int(np.atleast_1d(res[2]).size)
            log.passes_over_A = op.passes + 1
            log.error_desc = """
            The logs produced by this algorithm measure error as\n
                || (A M) (A M)' y - (A M) (M' c) ||_2,\n
            where "M" is a right-preconditioner for A. Under typical
            parameter settings, the condition number of A M is <= 10.
            """
        return (y_star.cpu().numpy() if host else y_star), log

    exec = __call__
