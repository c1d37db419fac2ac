"""Thin torch-tensor wrappers over the C ABI (``include/parla_b200.h``).

PyTorch is used for device memory, streams and (elsewhere) ``torch.distributed`` only; every
O(m n) operation below is a hand-written sm_100a kernel in ``libparla_b200.so``.  All wrappers
enqueue on the caller's current CUDA stream and never synchronise (unless documented).
"""
import ctypes

import torch

from . import _lib

PASS_DOT, PASS_AXPY, PASS_AXPY_G = 1, 2, 4
PASS_MAX_N = 8192
LSQR_NDOUBLE, LSQR_NINT = 32, 8
LSQR_SA = 15            # dstate[15:17] = (sa, su) of the next pass
F64 = torch.float64


PASS_TIMINGS = None     # set to a list to record (flags, m, n, start_event, end_event) per streaming pass


def launch_count():
    return int(_lib.load().pla_launch_count())


def note_launches(n):
    _lib.load().pla_note_launches(int(n))


def reserve_pass_workspace(device, n_max):
    """Allocate the streaming pass's scratch for the CURRENT stream now (so that a CUDA-graph capture on this
    stream finds it and allocates nothing)."""
    lib = _lib.load()
    return Workspace.get(device, lib.pla_stream_pass_workspace_bytes(1, int(n_max)), "pass")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _req(t, name, dtype=F64):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t


def _rowmajor(t, name):
    """Return (tensor, ld) for a 2-D row-major view (unit stride along columns)."""
    _req(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D")
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    if ld < t.shape[1]:
        t = t.contiguous()
        ld = t.shape[1]
    return t, ld


def _vec(t, name):
    _req(t, name)
    if t.dim() != 1:
        raise ValueError(f"{name} must be 1-D")
    return t if (t.numel() <= 1 or t.stride(0) == 1) else t.contiguous()


class Workspace:
    """Grow-only device scratch buffers handed to the C ABI as (ptr, bytes): one per (device, stream, tag), so
    work enqueued on different streams never shares partial-sum scratch.  A buffer that is outgrown goes back to
    torch's caching allocator, which is stream-ordered for the stream that allocated it -- the same stream
    that still has kernels using it -- so the hand-over is safe."""

    _bufs = {}

    @classmethod
    def get(cls, device, nbytes, tag="main"):
        device = torch.device(device)
        key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf


def num_sms():
    return _lib.load().pla_num_sms()


# ---------------------------------------------------------------------------------------- K4
def stream_pass(A, *, w=None, u=None, g=None, sc=None, sa=1.0, su=0.0, zss=None, flags=PASS_DOT, istop=None, comm=None):
    """One streaming read of A:  u <- sa*(A w) + su*u ;  z = A^T q ;  zss = [z, |u|^2].

    Returns zss (n+1 doubles).  See pla_stream_pass_f64.  Matrices wider than PASS_MAX_N columns go through
    column blocks (two reads of A when both products are requested: z needs the complete u).
    With ``comm`` (a parallel.PeerComm: A is this rank's row block) zss comes back summed over the ranks -- inside
    the reduce kernel, over NVLink peer memory (pla_stream_pass_peer_f64); the column-blocked path all-reduces.
    """
    lib = _lib.load()
    A, lda = _rowmajor(A, "A")
    m, n = A.shape
    if zss is None:
        zss = torch.empty(n + 1, dtype=F64, device=A.device)
    if n > PASS_MAX_N or (n % 2 == 1 and n > PASS_MAX_N // 2) or (comm is not None and n + 1 > comm.slot_lines):
        zss = _stream_pass_wide(A, w, u, g, sc, sa, su, zss, flags, istop)
        if comm is not None and comm.world > 1:
            torch.distributed.all_reduce(zss, group=comm.group)
        return zss
    nb = lib.pla_stream_pass_workspace_bytes(m, n)
    ws = Workspace.get(A.device, nb, "pass")
    rec = PASS_TIMINGS
    if rec is not None and torch.cuda.is_current_stream_capturing():
        rec = None
    if rec is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if comm is None:
        rc = lib.pla_stream_pass_f64(A.data_ptr(), m, n, lda, _p(w), _p(u), _p(g), _p(sc), float(sa), float(su),
                                     zss.data_ptr(), int(flags), _p(istop), ws.data_ptr(), ws.numel(), _stream())
    else:
        rc = lib.pla_stream_pass_peer_f64(A.data_ptr(), m, n, lda, _p(w), _p(u), _p(g), _p(sc), float(sa), float(su),
                                          zss.data_ptr(), int(flags), _p(istop), ws.data_ptr(), ws.numel(), comm.ptrs,
                                          comm.rank, comm.world, comm.slot_lines, comm.next_epoch(), _stream())
    _lib.check(rc, "pla_stream_pass_f64")
    if rec is not None:
        e1.record()
        rec.append((int(flags), m, n, e0, e1))
    return zss


FUSED_MAX_R, FUSED_MAX_NIN = 2048, 4096      # limits of pla_lsqr_fused_step_f64


def stream_pass_parts(A, *, w, u, sc=None, sa=1.0, su=0.0, flags=PASS_DOT | PASS_AXPY, istop=None):
    """The streaming pass WITHOUT its reduce launch (pla_stream_pass_parts_f64): returns (workspace, number of
    partials, offset of the |u|^2 partials) for :func:`lsqr_fused_step`.  Narrow / even-n shapes only (the caller
    checks: no column blocks, no peer exchange)."""
    lib = _lib.load()
    A, lda = _rowmajor(A, "A")
    m, n = A.shape
    ws = Workspace.get(A.device, lib.pla_stream_pass_workspace_bytes(m, n), "pass")
    parts = (ctypes.c_int64 * 2)()
    rc = lib.pla_stream_pass_parts_f64(A.data_ptr(), m, n, lda, _p(w), _p(u), None, _p(sc), float(sa), float(su),
                                       int(flags), _p(istop), ws.data_ptr(), ws.numel(),
                                       ctypes.addressof(parts), _stream())
    _lib.check(rc, "pla_stream_pass_parts_f64")
    return ws, int(parts[0]), int(parts[1])


def lsqr_fused_step(M, ws, nparts, ss_offset, zss, t, x, v, w, xw, dstate, istate, hist):
    """Reduce of the pass's partials, t = M^T z, the LSQR step and xw = M v_new in one cluster launch
    (pla_lsqr_fused_step_f64).  ``ws`` None / nparts 0: z and |u|^2 are taken from ``zss`` instead."""
    M, ldm = _rowmajor(M, "M")
    n_in, r = M.shape
    rc = _lib.load().pla_lsqr_fused_step_f64(n_in, r, M.data_ptr(), ldm, _p(ws), int(nparts), int(ss_offset),
                                             zss.data_ptr(), t.data_ptr(), x.data_ptr(), v.data_ptr(), w.data_ptr(),
                                             xw.data_ptr(), dstate.data_ptr(), istate.data_ptr(), hist.data_ptr(),
                                             _stream())
    _lib.check(rc, "pla_lsqr_fused_step_f64")


class FusedIteration:
    """The two launches of a fused LSQR iteration (:func:`stream_pass_parts` then :func:`lsqr_fused_step`) with their
    argument lists built ONCE: every buffer of the iteration is fixed, so the per-iteration host work is two ctypes
    calls instead of two wrappers (~17 us less per iteration, which matters when an iteration is ~60 us of GPU time).
    Calls made on another stream than the one it was built on (a CUDA-graph capture) go through the wrappers."""

    def __init__(self, A, M, xw, u, sc, istop, zss, t, x, v, w, dstate, istate, hist):
        self.lib = _lib.load()
        self.A, lda = _rowmajor(A, "A")
        self.M, ldm = _rowmajor(M, "M")
        m, n = self.A.shape
        n_in, r = self.M.shape
        self.stream = _stream()
        self.ws = Workspace.get(self.A.device, self.lib.pla_stream_pass_workspace_bytes(m, n), "pass")
        self.parts = (ctypes.c_int64 * 2)()
        self.keep = (xw, u, sc, istop, zss, t, x, v, w, dstate, istate, hist)       # keeps every buffer alive
        self.args_pass = (self.A.data_ptr(), m, n, lda, xw.data_ptr(), u.data_ptr(), None, _p(sc), 1.0, 0.0,
                          int(PASS_DOT | PASS_AXPY), _p(istop), self.ws.data_ptr(), self.ws.numel(),
                          ctypes.addressof(self.parts), self.stream)
        self.head_step = (n_in, r, self.M.data_ptr(), ldm, self.ws.data_ptr())
        self.tail_step = (zss.data_ptr(), t.data_ptr(), x.data_ptr(), v.data_ptr(), w.data_ptr(), xw.data_ptr(),
                          dstate.data_ptr(), istate.data_ptr(), hist.data_ptr(), self.stream)
        self.args_step = None

    def __call__(self):
        if _stream() != self.stream:
            xw, u, sc, istop, zss, t, x, v, w, dstate, istate, hist = self.keep
            ws, nparts, ss_off = stream_pass_parts(self.A, w=xw, u=u, sc=sc, istop=istop)
            lsqr_fused_step(self.M, ws, nparts, ss_off, zss, t, x, v, w, xw, dstate, istate, hist)
            return
        rc = self.lib.pla_stream_pass_parts_f64(*self.args_pass)
        if rc != 0:
            _lib.check(rc, "pla_stream_pass_parts_f64")
        if self.args_step is None:          # (the partial count of a shape never changes)
            self.args_step = self.head_step + (int(self.parts[0]), int(self.parts[1])) + self.tail_step
        rc = self.lib.pla_lsqr_fused_step_f64(*self.args_step)
        if rc != 0:
            _lib.check(rc, "pla_lsqr_fused_step_f64")


WIDE_BLOCK = 4096       # column-block width of the wide-matrix path (even, so every block keeps 16-byte aligned rows)


def _stream_pass_wide(A, w, u, g, sc, sa, su, zss, flags, istop):
    """n > PASS_MAX_N (or odd n > PASS_MAX_N / 2): the per-thread column ownership of the fused kernel does not
    cover the row, so the products are taken over column blocks A[:, c0:c1] (row tiles of a block are fetched with
    one bulk copy per row).  u <- sa * sum_k A_k w_k + su * u accumulates over the blocks; z_k = A_k^T q needs the
    finished u, hence a second sweep when both are requested -- the reference's own two-dgemv formulation
    (parla/comps/preconditioning.py:30,34).  After LSQR has stopped (``istop`` set) every launch is a no-op."""
    m, n = A.shape
    do_dot, do_axpy = bool(flags & PASS_DOT), bool(flags & PASS_AXPY)
    blocks = [(c0, min(c0 + WIDE_BLOCK, n)) for c0 in range(0, n, WIDE_BLOCK)]
    tmp = torch.empty(WIDE_BLOCK + 1, dtype=F64, device=A.device)
    if do_dot:
        sc_next = None
        if sc is not None:
            sc_next = torch.stack((sc[0], torch.ones((), dtype=F64, device=A.device)))
        for i, (c0, c1) in enumerate(blocks):
            first = i == 0
            stream_pass(A[:, c0:c1], w=w[c0:c1], u=u, sc=sc if first else sc_next, sa=sa, su=su if first else 1.0,
                        zss=tmp[:c1 - c0 + 1], flags=PASS_DOT, istop=istop)
        if not do_axpy:
            zss[n:n + 1].copy_(tmp[blocks[-1][1] - blocks[-1][0]:blocks[-1][1] - blocks[-1][0] + 1])
            return zss
    if do_axpy:
        fl = PASS_AXPY | (flags & PASS_AXPY_G)
        for c0, c1 in blocks:
            stream_pass(A[:, c0:c1], u=u, g=g, zss=tmp[:c1 - c0 + 1], flags=fl, istop=istop)
            zss[c0:c1].copy_(tmp[:c1 - c0])
        zss[n:n + 1].copy_(tmp[blocks[-1][1] - blocks[-1][0]:blocks[-1][1] - blocks[-1][0] + 1])
    return zss


def matvec(A, x, alpha=1.0, y=None, beta=0.0):
    """y <- alpha * A @ x + beta * y  (fresh y when None).  Returns (y, zss) with zss[-1] = |y|^2."""
    A2, _ = _rowmajor(A, "A")
    x = _vec(x, "x")
    if y is None:
        y = torch.zeros(A2.shape[0], dtype=F64, device=A2.device)
        beta = 0.0
    zss = stream_pass(A2, w=x, u=y, sa=alpha, su=beta, flags=PASS_DOT)
    return y, zss


def rmatvec(A, u):
    """z = A^T u.  Returns zss (z = zss[:n], zss[n] = |u|^2)."""
    A2, _ = _rowmajor(A, "A")
    return stream_pass(A2, u=_vec(u, "u"), flags=PASS_AXPY)


# ---------------------------------------------------------------------------------------- K3b
def trsv_upper(R, b, trans=False, out=None, istop=None):
    lib = _lib.load()
    _req(R, "R")
    n = R.shape[0]
    if R.stride(1) != 1:
        R = R.contiguous()
    b = _vec(b, "b")
    if out is None:
        out = torch.empty(n, dtype=F64, device=R.device)
    rc = lib.pla_trsv_upper_f64(R.data_ptr(), n, R.stride(0), 1 if trans else 0, b.data_ptr(), out.data_ptr(),
                                _p(istop), _stream())
    _lib.check(rc, "pla_trsv_upper_f64")
    return out


def trtri_upper(R):
    """Explicit inverse of an upper-triangular matrix (strict lower part of R ignored), blocked:
    32x32 diagonal blocks by back substitution, then X12 = -X11 (R12 X22) level by level with the
    DMMA GEMM.  Returns a dense row-major n x n tensor (zeros below the diagonal)."""
    lib = _lib.load()
    _req(R, "R")
    n = R.shape[0]
    if R.stride(1) != 1:
        R = R.contiguous()
    X = torch.zeros(n, n, dtype=F64, device=R.device)
    _lib.check(lib.pla_trtri_diag_f64(R.data_ptr(), n, R.stride(0), X.data_ptr(), n, _stream()), "pla_trtri_diag_f64")
    s = 32
    while s < n and s <= 64:          # small levels: every pair of the level in one launch
        _lib.check(lib.pla_trtri_merge_f64(R.data_ptr(), n, R.stride(0), X.data_ptr(), n, s, _stream()),
                   "pla_trtri_merge_f64")
        s *= 2
    while s < n:
        for i0 in range(0, n, 2 * s):
            i1, i2 = i0 + s, min(i0 + 2 * s, n)
            if i1 >= n:
                break
            T = gemm(R[i0:i1, i1:i2], X[i1:i2, i1:i2])                      # R12 X22
            gemm(X[i0:i1, i0:i1], T, alpha=-1.0, beta=0.0, out=X[i0:i1, i1:i2])   # X12 = -X11 T
        s *= 2
    return X


# ---------------------------------------------------------------------------------------- LSQR state
def lsqr_init(t, zss, bsq, atol, btol, conlim, iter_lim, x0, x, v, w, dstate, istate):
    rc = _lib.load().pla_lsqr_init_f64(x.numel(), t.data_ptr(), zss.data_ptr(), bsq.data_ptr(), float(atol),
                                       float(btol), float(conlim), int(iter_lim), _p(x0), x.data_ptr(),
                                       v.data_ptr(), w.data_ptr(), dstate.data_ptr(), istate.data_ptr(), _stream())
    _lib.check(rc, "pla_lsqr_init_f64")


def lsqr_step(t, zss, x, v, w, dstate, istate, hist):
    rc = _lib.load().pla_lsqr_step_f64(x.numel(), t.data_ptr(), zss.data_ptr(), x.data_ptr(), v.data_ptr(),
                                       w.data_ptr(), dstate.data_ptr(), istate.data_ptr(), hist.data_ptr(), _stream())
    _lib.check(rc, "pla_lsqr_step_f64")


def lsqr_ridge(sd, xw, ub, zss, sc=None, sa=1.0, su=0.0, istop=None):
    rc = _lib.load().pla_lsqr_ridge_f64(ub.numel(), float(sd), _p(xw), ub.data_ptr(), _p(sc), float(sa), float(su),
                                        zss.data_ptr(), _p(istop), _stream())
    _lib.check(rc, "pla_lsqr_ridge_f64")


def lsqr_under_init(cpc, u, atol, btol, conlim, iter_lim, dstate, istate):
    _lib.check(_lib.load().pla_lsqr_under_init_f64(u.numel(), cpc.data_ptr(), u.data_ptr(), float(atol), float(btol),
                                                   float(conlim), int(iter_lim), dstate.data_ptr(), istate.data_ptr(),
                                                   _stream()), "pla_lsqr_under_init_f64")


def lsqr_under_init2(nz, zss, dstate, istate):
    _lib.check(_lib.load().pla_lsqr_under_init2_f64(nz, zss.data_ptr(), dstate.data_ptr(), istate.data_ptr(), _stream()),
               "pla_lsqr_under_init2_f64")


def lsqr_under_head(t, u, dstate, istate):
    _lib.check(_lib.load().pla_lsqr_under_head_f64(u.numel(), t.data_ptr(), u.data_ptr(), dstate.data_ptr(),
                                                   istate.data_ptr(), _stream()), "pla_lsqr_under_head_f64")


def lsqr_under_tail(nz, zss, dstate, istate, hist):
    _lib.check(_lib.load().pla_lsqr_under_tail_f64(nz, zss.data_ptr(), dstate.data_ptr(), istate.data_ptr(),
                                                   hist.data_ptr(), _stream()), "pla_lsqr_under_tail_f64")


def lsqr_under_long(vt, x, w, dstate, istate, itn, add_to_ww=False):
    ws = Workspace.get(vt.device, 8 * 4 * 148 * 4, "lsqr_long")
    _lib.check(_lib.load().pla_lsqr_under_long_f64(vt.numel(), vt.data_ptr(), x.data_ptr(), w.data_ptr(),
                                                   dstate.data_ptr(), istate.data_ptr(), int(itn), 1 if add_to_ww else 0,
                                                   ws.data_ptr(), ws.numel(), _stream()), "pla_lsqr_under_long_f64")


# ---------------------------------------------------------------------------------------- PCG state
PCG_ERR = 1             # dstate slot of |r| (see include/parla_b200.h)


def pcg_residual(rhs, gx, delta, x, r, dstate, istate, init=False, tol=0.0):
    _lib.check(_lib.load().pla_pcg_residual_f64(rhs.numel(), rhs.data_ptr(), _p(gx), float(delta), _p(x), r.data_ptr(),
                                                dstate.data_ptr(), istate.data_ptr(), 1 if init else 0, float(tol),
                                                _stream()), "pla_pcg_residual_f64")


def pcg_direction(r, s, p, dstate, istate, init=False, iter_lim=0):
    _lib.check(_lib.load().pla_pcg_direction_f64(r.numel(), r.data_ptr(), s.data_ptr(), p.data_ptr(), dstate.data_ptr(),
                                                 istate.data_ptr(), 1 if init else 0, int(iter_lim), _stream()),
               "pla_pcg_direction_f64")


def pcg_update(gp, delta, p, x, r, dstate, istate, hist, recompute):
    _lib.check(_lib.load().pla_pcg_update_f64(p.numel(), gp.data_ptr(), float(delta), p.data_ptr(), x.data_ptr(),
                                              r.data_ptr(), dstate.data_ptr(), istate.data_ptr(), hist.data_ptr(),
                                              1 if recompute else 0, _stream()), "pla_pcg_update_f64")


def sumsq(x, out=None):
    lib = _lib.load()
    x = _vec(x, "x")
    if out is None:
        out = torch.empty(1, dtype=F64, device=x.device)
    ws = Workspace.get(x.device, lib.pla_sumsq_workspace_bytes(x.numel()), "sumsq")
    _lib.check(lib.pla_sumsq_f64(x.data_ptr(), x.numel(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
               "pla_sumsq_f64")
    return out


# ---------------------------------------------------------------------------------------- GEMM
def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, out=None):
    """out = alpha * op(A) @ op(B) + beta * out   (FP64 DMMA kernel)."""
    lib = _lib.load()
    A, lda = _rowmajor(A, "A")
    B, ldb = _rowmajor(B, "B")
    M, K = (A.shape[1], A.shape[0]) if transa else A.shape
    Kb, N = (B.shape[1], B.shape[0]) if transb else B.shape
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions differ ({K} vs {Kb})")
    if out is None:
        out = torch.empty(M, N, dtype=F64, device=A.device)
        beta = 0.0
    elif out.shape != (M, N) or out.stride(1) != 1:
        raise ValueError("gemm: out must be a row-major (M, N) tensor")
    _req(out, "out")
    nb = lib.pla_gemm_workspace_bytes(M, N, K)
    ws = Workspace.get(A.device, nb, "gemm")
    ldc = out.stride(0) if M > 1 else max(N, 1)
    rc = lib.pla_gemm_f64(int(transa), int(transb), M, N, K, float(alpha), A.data_ptr(), lda, B.data_ptr(), ldb,
                          float(beta), out.data_ptr(), ldc, ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "pla_gemm_f64")
    return out


# ---------------------------------------------------------------------------------------- K1
def sketch_gauss(A, d, seed, scale, out, bvec=None, col_offset=0, beta=0.0):
    """out[:d, :n(+1)] = beta*out + scale * G(seed)[:, col_offset:col_offset+m] @ [A | bvec]."""
    lib = _lib.load()
    A, lda = _rowmajor(A, "A")
    m, n = A.shape
    nb = lib.pla_sketch_gauss_workspace_bytes(d, n, m)
    ws = Workspace.get(A.device, nb, "gemm")
    rc = lib.pla_sketch_gauss_f64(A.data_ptr(), m, n, lda, _p(bvec), d, ctypes.c_uint64(seed), col_offset,
                                  float(scale), float(beta), out.data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(),
                                  _stream())
    _lib.check(rc, "pla_sketch_gauss_f64")
    return out


def philox_normal_fill(rows, cols, seed, scale=1.0, row_offset=0, col_offset=0, device="cuda"):
    out = torch.empty(rows, cols, dtype=F64, device=device)
    rc = _lib.load().pla_philox_normal_fill_f64(out.data_ptr(), rows, cols, cols, ctypes.c_uint64(seed), row_offset,
                                                col_offset, float(scale), _stream())
    _lib.check(rc, "pla_philox_normal_fill_f64")
    return out


# ---------------------------------------------------------------------------------------- K2
class SjltPlan:
    """Destination-major plan of a d x m SJLT given in index form (rows[m,k] int32, signs[m,k] int8)."""

    def __init__(self, rows, signs, d, validate=False):
        lib = _lib.load()
        _req(rows, "rows", torch.int32)
        _req(signs, "signs", torch.int8)
        rows, signs = rows.contiguous(), signs.contiguous()
        self.m, self.k = rows.shape
        self.d = int(d)
        self.buf = torch.empty(lib.pla_sjlt_plan_bytes(self.d, self.m, self.k), dtype=torch.uint8, device=rows.device)
        ws = Workspace.get(rows.device, lib.pla_sjlt_plan_workspace_bytes(self.d, self.m, self.k), "sjlt")
        rc = lib.pla_sjlt_plan_f64(rows.data_ptr(), signs.data_ptr(), self.m, self.k, self.d, self.buf.data_ptr(),
                                   ws.data_ptr(), ws.numel(), _stream())
        _lib.check(rc, "pla_sjlt_plan_f64")
        if validate:
            bad = ctypes.c_int64(0)
            _lib.check(lib.pla_sjlt_plan_status(self.buf.data_ptr(), ctypes.byref(bad)), "pla_sjlt_plan_status")
            if bad.value:
                raise ValueError(f"SJLT index form has {bad.value} row indices outside [0, {self.d})")

    def apply(self, A, scale, out, bvec=None, out_b=None, accumulate=False):
        lib = _lib.load()
        A, lda = _rowmajor(A, "A")
        m, n = A.shape
        if m != self.m:
            raise ValueError(f"SJLT has {self.m} columns but A has {m} rows")
        ldob = out_b.stride(0) if (out_b is not None and out_b.numel() > 1) else 1
        nb = lib.pla_sjlt_apply_workspace_bytes(self.d, n)
        ws = Workspace.get(A.device, nb, "sjlt_apply") if nb else None
        rc = lib.pla_sjlt_apply_f64(self.buf.data_ptr(), self.d, self.m, self.k, A.data_ptr(), n, lda, _p(bvec),
                                    float(scale), out.data_ptr(), out.stride(0), _p(out_b), ldob,
                                    1 if accumulate else 0, _p(ws), ws.numel() if ws is not None else 0, _stream())
        _lib.check(rc, "pla_sjlt_apply_f64")
        return out


def sjlt_rmatvec(rows, signs, d, v, scale, out=None):
    """out = scale * S^T v for the SJLT in index form (rows/signs [m, k])."""
    _req(rows, "rows", torch.int32)
    _req(signs, "signs", torch.int8)
    rows, signs = rows.contiguous(), signs.contiguous()
    v = _vec(v, "v")
    if v.numel() != d:
        raise ValueError(f"S^T v: v has {v.numel()} entries, the operator has {d} rows")
    m, k = rows.shape
    if out is None:
        out = torch.empty(m, dtype=F64, device=v.device)
    _lib.check(_lib.load().pla_sjlt_rmatvec_f64(rows.data_ptr(), signs.data_ptr(), m, k, int(d), v.data_ptr(),
                                                float(scale), out.data_ptr(), _stream()), "pla_sjlt_rmatvec_f64")
    return out


def gauss_rmatvec(d, m, seed, scale, v, col_offset=0, out=None):
    """out = scale * G(seed)[0:d, col_offset:col_offset+m]^T v  (virtual Gaussian operator)."""
    v = _vec(v, "v")
    if v.numel() != d:
        raise ValueError(f"S^T v: v has {v.numel()} entries, the operator has {d} rows")
    if out is None:
        out = torch.empty(m, dtype=F64, device=v.device)
    _lib.check(_lib.load().pla_gauss_rmatvec_f64(int(d), int(m), ctypes.c_uint64(seed), int(col_offset), float(scale),
                                                 v.data_ptr(), out.data_ptr(), _stream()), "pla_gauss_rmatvec_f64")
    return out


def sjlt_generate(d, m, k, seed, col_offset=0, device="cuda"):
    rows = torch.empty(m, k, dtype=torch.int32, device=device)
    signs = torch.empty(m, k, dtype=torch.int8, device=device)
    rc = _lib.load().pla_sjlt_generate(d, m, k, ctypes.c_uint64(seed), col_offset, rows.data_ptr(), signs.data_ptr(),
                                       _stream())
    _lib.check(rc, "pla_sjlt_generate")
    return rows, signs


# ---------------------------------------------------------------------------------------- SRCT
def srct_weights(k, m, j0, ncols, jmap=None, e=None, sgn=None, with_sin=False, out=None):
    """Cosine (and sine) weights of the orthonormal DCT-II for the frequencies ``k`` (int64 device vector) and
    ``ncols`` positions j = jmap[j0 + c] (or j0 + c); see pla_srct_weights_f64."""
    _req(k, "k", torch.int64)
    g = k.numel()
    width = (2 if with_sin else 1) * ncols
    if out is None:
        out = torch.empty(g, width, dtype=F64, device=k.device)
    _lib.check(_lib.load().pla_srct_weights_f64(k.data_ptr(), g, int(m), int(j0), int(ncols), _p(jmap), _p(e), _p(sgn),
                                                1 if with_sin else 0, out.data_ptr(), out.stride(0), _stream()),
               "pla_srct_weights_f64")
    return out


def gather_rows_scale(A, perm, e, c0, nb, out):
    """out[t, :nb] = e[t] * A[perm[t], c0:c0+nb]."""
    A, lda = _rowmajor(A, "A")
    rows = out.shape[0]
    _lib.check(_lib.load().pla_gather_rows_scale_f64(A.data_ptr(), lda, _p(perm), _p(e), rows, int(c0), int(nb),
                                                     out.data_ptr(), out.stride(0), _stream()),
               "pla_gather_rows_scale_f64")
    return out


# ---------------------------------------------------------------------------------------- K3a / K6
def geqrf(W, ncols_factor):
    """In-place Householder QR of the leading columns of the row-major W; returns tau."""
    lib = _lib.load()
    _req(W, "W")
    if W.dim() != 2 or W.stride(1) != 1:
        raise ValueError("geqrf: W must be a row-major 2-D tensor")
    M, N = W.shape
    tau = torch.empty(ncols_factor, dtype=F64, device=W.device)
    ws = Workspace.get(W.device, lib.pla_qr_workspace_bytes(M, N), "qr")
    rc = lib.pla_geqrf_f64(W.data_ptr(), M, N, W.stride(0), ncols_factor, tau.data_ptr(), ws.data_ptr(), ws.numel(),
                           _stream())
    _lib.check(rc, "pla_geqrf_f64")
    return tau


QR_BLOCK = 128          # column block of the Householder QR (QR_NBO in csrc/qr.cu)


def qr_block_workspace(device, M, n_layout):
    lib = _lib.load()
    return Workspace.get(device, lib.pla_qr_workspace_bytes(M, n_layout), "qr")


def qr_factor_block(P, r0, c0, jb, tau_blk, block_index, n_layout, ws):
    """Factor columns [c0, c0 + jb) of the row-major P, rows [r0, M), in place (see pla_qr_factor_block_f64)."""
    _req(P, "P")
    rc = _lib.load().pla_qr_factor_block_f64(P.data_ptr(), P.shape[0], P.stride(0), int(r0), int(c0), int(jb),
                                             tau_blk.data_ptr(), int(block_index), int(n_layout), ws.data_ptr(),
                                             ws.numel(), _stream())
    _lib.check(rc, "pla_qr_factor_block_f64")


def qr_apply_block(M, r0, jb, tau_blk, C, n_layout, ws):
    """C[r0:, :] <- Q_block^T C[r0:, :] with the reflectors of the last qr_factor_block call (same workspace)."""
    nc = C.shape[1]
    if nc == 0:
        return
    _req(C, "C")
    rc = _lib.load().pla_qr_apply_block_f64(int(M), int(r0), int(jb), tau_blk.data_ptr(), C.data_ptr(), C.stride(0), nc,
                                            int(n_layout), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "pla_qr_apply_block_f64")


def orgqr(W, tau, K=None):
    lib = _lib.load()
    M = W.shape[0]
    K = tau.numel() if K is None else K
    Q = torch.empty(M, K, dtype=F64, device=W.device)
    ws = Workspace.get(W.device, lib.pla_qr_workspace_bytes(M, K), "qr")
    rc = lib.pla_orgqr_f64(W.data_ptr(), M, K, W.stride(0), tau.data_ptr(), Q.data_ptr(), K, ws.data_ptr(),
                           ws.numel(), _stream())
    _lib.check(rc, "pla_orgqr_f64")
    return Q


def qr_economic(Y):
    """(Q, R) of a tall row-major matrix, LAPACK sign conventions (scipy.linalg.qr(mode='economic'))."""
    Y, _ = _rowmajor(Y, "Y")
    W = Y.clone()
    M, N = W.shape
    K = min(M, N)
    tau = geqrf(W, K)
    Q = orgqr(W, tau, K)
    R = torch.triu(W[:K, :])
    return Q, R
