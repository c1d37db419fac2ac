"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the reference itself.

Run in the BUILD container (the reference is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py [/root/reference]

For every case it (1) runs the unmodified reference (imported from the path given), (2) runs
``oracle/parla_oracle.py`` on the same inputs and seeds, (3) asserts the two agree to round-off
(the numbers printed), and (4) stores the REFERENCE's outputs plus whatever is needed to rebuild
the inputs (seeds; the sketching operator in index form) as a small fixture.  This is what pins
the oracle (see the header of parla_oracle.py).
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = next((a for a in sys.argv[1:] if not a.startswith("--")), "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
np.NaN = np.nan            # numpy-2 shim needed by the reference's QB2 (qb.py:468)
warnings.filterwarnings("ignore")

import parla as rla                                             # noqa: E402  (the reference)
import parla.comps.sketchers.oblivious as rsko                  # noqa: E402
import parla.utils.linalg_wrappers as rulaw                     # noqa: E402
import parla.tests.matmakers as rmm                             # noqa: E402
from oracle import parla_oracle as orc                          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class Tape:
    """Wraps a sketch-operator generator and records every operator it returns."""

    def __init__(self, gen):
        self.gen, self.ops = gen, []

    def __call__(self, n_rows, n_cols, rng):
        S = self.gen(n_rows, n_cols, rng)
        self.ops.append(S)
        return S


def lsq_problem(m, n, seed, cond=1.0):
    """Gaussian A with optional column scaling, b = A x0 + 0.1 noise (SURVEY.md 8d, cfg1 recipe)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    if cond != 1.0:
        A = A * np.logspace(0, np.log10(cond), n)
    x0 = rng.standard_normal(n)
    b = A @ x0 + 0.1 * rng.standard_normal(m)
    return A, b


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def spo_case(name, m, n, seed, sketch, mode, delta, sf=4, tol=1e-12, iter_lim=100, cond=1.0,
             store_S=True, rng_seed=1, big=False):
    A, b = lsq_problem(m, n, seed, cond)
    ref_gen = Tape({'sjlt': rsko.SkOpSJ(8), 'gauss': rsko.SkOpGA(), 'srct': rsko.SkOpTC()}[sketch])
    orc_gen = Tape({'sjlt': orc.SkOpSJ(8), 'gauss': orc.SkOpGA(), 'srct': orc.SkOpTC()}[sketch])
    x_ref, log_ref = rla.SPO(ref_gen, sf, mode)(A, b, delta, tol, iter_lim,
                                               np.random.default_rng(rng_seed), logging=True)
    x_orc, log_orc = orc.SPO(orc_gen, sf, mode)(A, b, delta, tol, iter_lim,
                                               np.random.default_rng(rng_seed), logging=True)
    S_ref, S_orc = ref_gen.ops[0], orc_gen.ops[0]
    if sketch == 'sjlt':
        same_S = (S_ref != S_orc).nnz == 0
        rows, signs, k = orc.sjlt_index_form(S_ref)
    elif sketch == 'srct':
        same_S = all(np.array_equal(a, c) for a, c in zip(S_ref.sketch_data, S_orc.sketch_data))
    else:
        same_S = bool(np.array_equal(S_ref, S_orc))
    e_x = relerr(x_orc, x_ref)
    kk = min(log_orc.errors.size, log_ref.errors.size)
    e_err = relerr(log_orc.errors[:kk], log_ref.errors[:kk])
    its = (log_ref.errors.size - 1, log_orc.errors.size - 1)
    print(f"{name:28s} iters ref/orc {its}  |dx|/|x| {e_x:.2e}  d(errors) {e_err:.2e}  S identical {same_S}")
    assert same_S, "oracle sketch operator differs from the reference's"
    assert e_x < 1e-11 and abs(its[0] - its[1]) <= 1 and e_err < 1e-6
    r = A @ x_ref - b
    fx = dict(m=m, n=n, seed=seed, cond=cond, sketch=sketch, mode=mode, delta=delta, sf=sf, tol=tol,
              iter_lim=iter_lim, rng_seed=rng_seed, x=x_ref, errors=log_ref.errors,
              resid_norm=np.linalg.norm(r), A_sha=digest(A), b_sha=digest(b))
    if big:
        # b = A @ x0 + noise goes through a BLAS gemv whose summation order depends on the host; a problem this
        # large is pinned by the hash of A (pure RNG output) and a probe of b instead of b's hash
        del fx["b_sha"]
        fx.update(b_probe=b[::4096].copy(), b_norm=np.linalg.norm(b))
    if sketch == 'sjlt':
        fx.update(S_sha=digest(rows) + digest(signs), vec_nnz=k)
        if store_S:
            fx.update(S_rows=rows.astype(np.int16 if rows.max() < 32768 else np.int32), S_signs=signs)
    elif sketch == 'srct':
        r, e, perm = S_ref.sketch_data
        fx.update(S_r=r.astype(np.int32), S_e_sign=np.sign(e).astype(np.int8), S_e_scale=np.abs(e[0]),
                  S_perm=perm.astype(np.int32))
    else:
        fx.update(S_sha=digest(S_ref))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)


def spo_rankdef_case(name, rng_seed, iter_lim=100, tol=1e-12):
    """The reference's `consistent_lowrank` problem (test_overdet_least_squares.py:23-34; rank 5 of 10 columns)
    through SPO(SkOpGA, 3, 'svd') (:408-411, with the tolerance the test meant: 1e-12).  Seed 1: the presolve is
    exact and accepted; seed 4: R is non-square and the presolve is off, so LSQR starts from the origin
    (least_squares.py:348-351)."""
    import parla.utils.sketching as rusk
    rng = np.random.default_rng(8923890298)
    m, n, rank = 100, 10, 5
    U = rusk.orthonormal_operator(m, rank, rng)
    sv = rng.random(rank) + 1e-4
    Vt = rusk.orthonormal_operator(rank, n, rng)
    A = (U * sv) @ Vt
    x0 = rng.standard_normal(n)
    b = A @ x0
    A2, b2, x02, Vt2 = orc.consistent_lowrank_problem()
    assert relerr(A2, A) < 1e-14 and relerr(b2, b) < 1e-14
    ref_gen, orc_gen = Tape(rsko.SkOpGA()), Tape(orc.SkOpGA())
    x_ref, log_ref = rla.SPO(ref_gen, 3, 'svd')(A, b, 0.0, tol, iter_lim, np.random.default_rng(rng_seed), logging=True)
    x_orc, log_orc = orc.SPO(orc_gen, 3, 'svd')(A, b, 0.0, tol, iter_lim, np.random.default_rng(rng_seed), logging=True)
    assert np.array_equal(ref_gen.ops[0], orc_gen.ops[0])
    e_x = relerr(x_orc, x_ref)
    print(f"{name:28s} iters ref/orc {(log_ref.errors.size - 1, log_orc.errors.size - 1)}  |dx|/|x| {e_x:.2e}  "
          f"|Ax-b| {np.linalg.norm(A @ x_ref - b):.1e}  min-norm gap {abs(np.linalg.norm(x_ref) - np.linalg.norm(Vt.T @ (Vt @ x0))):.1e}")
    assert e_x < 1e-10 and log_ref.errors.size == log_orc.errors.size
    np.savez_compressed(os.path.join(OUT, name + ".npz"), rng_seed=rng_seed, iter_lim=iter_lim, tol=tol, x=x_ref,
                        errors=log_ref.errors, resid_norm=np.linalg.norm(A @ x_ref - b), A=A, b=b, S=ref_gen.ops[0],
                        x_minnorm=Vt.T @ (Vt @ x0))


def srct_cases():
    """SPO with the SRCT operator (test_overdet_least_squares.py:265-272,312-319,360-367), incl. a prime row
    count (no usable two-level factorisation) and a wider problem."""
    spo_case("spo_srct_qr_600x40", 600, 40, 11, 'srct', 'qr', 0.0)
    spo_case("spo_srct_qr_ridge_600x40", 600, 40, 12, 'srct', 'qr', 0.25, sf=2)
    spo_case("spo_srct_svd_600x40", 600, 40, 11, 'srct', 'svd', 0.0)
    spo_case("spo_srct_chol_600x40", 600, 40, 11, 'srct', 'chol', 0.0, sf=2)
    spo_case("spo_srct_qr_prime_1009x33", 1009, 33, 16, 'srct', 'qr', 0.0)
    spo_case("spo_srct_qr_cond1e4_4096x96", 4096, 96, 17, 'srct', 'qr', 0.0, cond=1e4)


def spu_case(name, m, n, seed, cond=1e3, sf=4, tol=1e-12, iter_lim=100, rng_seed=1):
    """Under-determined least squares (min |y| s.t. A'y = c), SPU1 (least_squares.py:425-494)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n)) * np.logspace(0, np.log10(cond), n)
    c = rng.standard_normal(n)
    ref_gen, orc_gen = Tape(rsko.SkOpSJ(8)), Tape(orc.SkOpSJ(8))
    y_ref, log_ref = rla.SPU1(ref_gen, sf)(A, c, tol, iter_lim, np.random.default_rng(rng_seed), logging=True)
    y_orc, log_orc = orc.SPU1(orc_gen, sf)(A, c, tol, iter_lim, np.random.default_rng(rng_seed), logging=True)
    assert (ref_gen.ops[0] != orc_gen.ops[0]).nnz == 0
    rows, signs, k = orc.sjlt_index_form(ref_gen.ops[0])
    e_y = relerr(y_orc, y_ref)
    its = (log_ref.errors.size - 1, log_orc.errors.size - 1)
    print(f"{name:28s} iters ref/orc {its}  |dy|/|y| {e_y:.2e}  constraint |A'y-c|/|c| "
          f"{np.linalg.norm(A.T @ y_ref - c) / np.linalg.norm(c):.2e}")
    assert e_y < 1e-9 and abs(its[0] - its[1]) <= 1, (e_y, its)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), m=m, n=n, seed=seed, cond=cond, sf=sf, tol=tol,
                        iter_lim=iter_lim, rng_seed=rng_seed, y_norm=np.linalg.norm(y_ref),
                        y_probe=y_ref[::max(1, m // 64)], errors=log_ref.errors, A_sha=digest(A), c_sha=digest(c),
                        S_rows=rows.astype(np.int16), S_signs=signs, vec_nnz=k)


def saddle_spectrum(kind, n, cond):
    if kind == 'lin':
        return np.linspace(cond ** 0.5, cond ** -0.5, num=n)
    return np.logspace(np.log10(cond) / 2, -np.log10(cond) / 2, num=n)


def sps_case(name, alg, m, n, cond, kind, delta, tol, iter_lim, seed=0, rng_seed=1, rhs_scale=1.0, sf=None):
    """Saddle-point drivers (saddlesys.py): alg in {'sps1', 'sps1_left', 'sps1_right', 'sps2'} on the
    reference's own test problem (test_saddlesys.py:11-57)."""
    import types
    import scipy.linalg as la
    import parla.drivers.saddlesys as rss
    import parla.comps.determiter.saddle as rdsad
    import parla.tests.test_drivers.test_optim.test_saddlesys as rts

    def solve(a, b, sym_pos=False, **kw):                      # scipy >= 1.11 dropped sym_pos
        return la.solve(a, b, assume_a='pos' if sym_pos else 'gen', **kw)
    rts.la = types.SimpleNamespace(solve=solve, norm=la.norm, lstsq=la.lstsq, LinAlgError=la.LinAlgError)

    spec = saddle_spectrum(kind, n, cond)
    ath = rts.make_simple_prob(m, n, spec, delta, np.random.default_rng(seed), rhs_scale=rhs_scale)
    A, b, c, x_opt, y_opt = orc.saddle_problem(m, n, spec, delta, np.random.default_rng(seed), rhs_scale)
    assert relerr(A, ath.A) < 1e-13 and relerr(b, ath.b) < 1e-13 and relerr(c, ath.c) < 1e-13
    # both arms get the SAME inputs: the oracle-built problem (equal to the reference's up to the last bit of
    # a BLAS nrm2; the digests in the fixture are of these arrays)
    gauss = alg in ('sps1_left', 'sps1_right')
    sf = sf if sf is not None else (0.85 if gauss else 3)
    ref_gen = Tape(rsko.SkOpGA() if gauss else rsko.SkOpSJ())
    orc_gen = Tape(orc.SkOpGA() if gauss else orc.SkOpSJ())
    if alg == 'sps2':
        ref_alg, orc_alg = rss.SPS2(ref_gen, sf, rdsad.PcSS2()), orc.SPS2(orc_gen, sf)
    else:
        ref_alg, orc_alg = rss.SPS1(ref_gen, sf, rdsad.PcSS1()), orc.SPS1(orc_gen, sf)
        if gauss:
            ref_alg.nystrom_strategy = orc_alg.nystrom_strategy = alg.split('_')[1]
    x_ref, y_ref, log_ref = ref_alg(A.copy(), b.copy(), c.copy(), delta, tol, iter_lim,
                                    np.random.default_rng(rng_seed), logging=True)
    x_orc, y_orc, log_orc = orc_alg(A.copy(), b.copy(), c.copy(), delta, tol, iter_lim,
                                    np.random.default_rng(rng_seed), logging=True)
    S_ref, S_orc = ref_gen.ops[0], orc_gen.ops[0]
    same_S = bool(np.array_equal(S_ref, S_orc)) if gauss else (S_ref != S_orc).nnz == 0
    e_x, e_y = relerr(x_orc, x_ref), relerr(y_orc, y_ref)
    its = (log_ref.errors.size - 1, log_orc.errors.size - 1)
    e_opt = relerr(x_ref, x_opt)
    print(f"{name:34s} iters ref/orc {its}  |dx|/|x| {e_x:.2e} |dy|/|y| {e_y:.2e}  ref vs x_opt {e_opt:.1e}  "
          f"S identical {same_S}")
    assert same_S and abs(its[0] - its[1]) <= 1 and e_x < 1e-8 and e_y < 1e-8, (e_x, e_y, its)
    fx = dict(alg=alg, m=m, n=n, cond=cond, kind=kind, delta=delta, tol=tol, iter_lim=iter_lim, seed=seed,
              rng_seed=rng_seed, rhs_scale=rhs_scale, sf=sf, x=x_ref, y_norm=np.linalg.norm(y_ref),
              y_probe=y_ref[::max(1, m // 64)], errors=log_ref.errors, x_opt=x_opt,
              A_sha=digest(A), b_sha=digest(b), c_sha=digest(c))
    if gauss:
        fx.update(S_sha=digest(S_ref))
    else:
        rows, signs, k = orc.sjlt_index_form(S_ref)
        fx.update(S_rows=rows.astype(np.int16), S_signs=signs, vec_nnz=k)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)


def sps_cases():
    sps_case("sps1_lin_1000x100", 'sps1', 1000, 100, 1e5, 'lin', 0.0, 1e-6 / 5e3, 50)
    sps_case("sps1_lin_ridge_1000x100", 'sps1', 1000, 100, 1e5, 'lin', 0.5, 1e-6 / 5e3, 50)
    sps_case("sps1_log_ridge_1000x100", 'sps1', 1000, 100, 1e5, 'log', 0.5, 1e-6 / 5e3, 50, rng_seed=4)
    sps_case("sps1_hiacc_500x50", 'sps1', 500, 50, 1e3, 'lin', 1.0, 1e-12, 50, rng_seed=15)
    sps_case("sps1_tiny_500x50", 'sps1', 500, 50, 1e8, 'lin', 0.0, 1e-8, 50, rhs_scale=1e-9, rng_seed=31)
    sps_case("sps1_nysleft_1000x100", 'sps1_left', 1000, 100, 1e5, 'lin', 0.0, 1e-10, 100)
    sps_case("sps1_nysleft_ridge_1000x100", 'sps1_left', 1000, 100, 1e5, 'log', 0.3, 1e-10, 100, rng_seed=42)
    sps_case("sps1_nysright_1000x100", 'sps1_right', 1000, 100, 1e5, 'lin', 0.0, 1e-10, 100)
    sps_case("sps1_nysright_ridge_1000x100", 'sps1_right', 1000, 100, 1e5, 'log', 0.3, 1e-10, 100, rng_seed=4)
    sps_case("sps2_lin_1000x100", 'sps2', 1000, 100, 1e5, 'lin', 0.0, 1e-12, 50)
    sps_case("sps2_lin_ridge_1000x100", 'sps2', 1000, 100, 1e5, 'lin', 0.5, 1e-12, 50)
    sps_case("sps2_log_ridge_1000x100", 'sps2', 1000, 100, 1e5, 'log', 0.5, 1e-12, 50, rng_seed=15)
    sps_case("sps2_hiacc_500x50", 'sps2', 500, 50, 1e3, 'lin', 1.0, 1e-12, 50, rng_seed=42)


def lowrank_case(name, m, n, rank, k, seed, blk=None, tol=np.nan, over=0, num_pass=2, evd=False, big=False,
                 spectrum_param=2.0):
    rng = np.random.default_rng(seed)
    if evd:
        B0 = rmm.rand_low_rank(m, rank, rank, rng)
        A = B0 @ B0.T
        A = 0.5 * (A + A.T)
        n = m
    else:
        A = rmm.exponent_spectrum(m, n, rank, rng, spectrum_param)
    A_orc = orc.exponent_spectrum(m, n, rank, np.random.default_rng(seed), spectrum_param) if not evd else A
    assert relerr(A_orc, A) < 1e-13

    def build(lib, sko, orth):
        rs = lib.RS1(sko, num_pass, orth, 1)
        rf = lib.RF1(rs)
        qb = lib.QB1(rf) if blk is None else lib.QB2(rf, blk, False)
        return (lib.EVD1(qb) if evd else lib.SVD1(qb)), qb

    ref_alg, ref_qb = build(rla, rsko.SkOpGA(), rulaw.orth)
    orc_alg, orc_qb = build(orc, orc.SkOpGA(), orc.orth)
    Q_ref, B_ref = ref_qb(A, k + over, tol, np.random.default_rng(7))
    Q_orc, B_orc = orc_qb(A, k + over, tol, np.random.default_rng(7))
    e_qb = relerr(Q_orc @ B_orc, Q_ref @ B_ref)
    out_ref = ref_alg(A, k, tol, over, np.random.default_rng(7))
    out_orc = orc_alg(A, k, tol, over, np.random.default_rng(7))
    if evd:
        V, lam = out_ref
        Vo, lamo = out_orc
        approx_ref, approx_orc = (V * lam) @ V.T, (Vo * lamo) @ Vo.T
        spec_ref, spec_orc = lam, lamo
    else:
        U, s, Vh = out_ref
        Uo, so, Vho = out_orc
        approx_ref, approx_orc = (U * s) @ Vh, (Uo * so) @ Vho
        spec_ref, spec_orc = s, so
    e_ap = relerr(approx_orc, approx_ref)
    e_sp = float(np.max(np.abs(spec_orc - spec_ref)) / np.max(np.abs(spec_ref)))
    print(f"{name:28s} QB cols {Q_ref.shape[1]}  d(QB) {e_qb:.2e}  d(approx) {e_ap:.2e}  d(spec) {e_sp:.2e}")
    assert Q_ref.shape == Q_orc.shape and e_qb < 1e-10 and e_ap < 1e-10 and e_sp < 1e-12
    extra = {}
    if big:
        # A is built with a LAPACK QR and GEMMs: bit-reproducible across hosts only for small shapes, so a large
        # matrix is pinned by a probe and its norm (to 1e-13) instead of a hash
        extra = dict(A_probe=A[::max(1, m // 16), ::max(1, n // 16)].copy(), A_fro=np.linalg.norm(A),
                     spectrum_param=spectrum_param)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), m=m, n=n, rank=rank, k=k, seed=seed,
                        blk=-1 if blk is None else blk, tol=tol, over=over, num_pass=num_pass,
                        evd=evd, spec=spec_ref, qb_cols=Q_ref.shape[1],
                        approx_fro=np.linalg.norm(approx_ref),
                        err_fro=np.linalg.norm(A - approx_ref), A_sha="" if big else digest(A),
                        approx_probe=approx_ref[::max(1, m // 16), ::max(1, n // 16)], **extra)


def qb3_evd2_cases():
    """QB3 (comps/qb.py:484-598; test_qb.py:414-489) and EVD2 (drivers/evd.py:290-381; test_evd.py:159-183)."""
    import parla.comps.qb as rqb
    import parla.drivers.evd as revd
    import parla.comps.sketchers.aware as raw
    for name, m, n, rank, k, blk, tol, seed in (("qb3_200x50", 200, 50, 30, 28, 4, np.nan, 41),
                                                ("qb3_tol_200x50", 200, 50, 30, 28, 4, 2e-2, 42),
                                                ("qb3_wide_60x240", 60, 240, 25, 24, 5, np.nan, 43)):
        A = orc.exponent_spectrum(m, n, rank, np.random.default_rng(seed), 3.0)
        Qr, Br = rqb.QB3(raw.RS1(rsko.SkOpGA(), 0, rulaw.orth, 1), blk)(A, k, tol, np.random.default_rng(7))
        Qo, Bo = orc.QB3(orc.RS1(orc.SkOpGA(), 0, orc.orth, 1), blk)(A, k, tol, np.random.default_rng(7))
        ap = Qr @ Br
        print(f"{name:28s} QB3 cols {Qr.shape[1]}  d(QB) {relerr(Qo @ Bo, ap):.2e}  |A-QB|/|A| {relerr(ap, A):.2e}")
        assert Qr.shape == Qo.shape and relerr(Qo @ Bo, ap) < 1e-10
        np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="qb3", m=m, n=n, rank=rank, k=k, blk=blk, tol=tol,
                            seed=seed, qb_cols=Qr.shape[1], approx_fro=np.linalg.norm(ap),
                            err_fro=np.linalg.norm(A - ap), A_sha=digest(A),
                            approx_probe=ap[::max(1, m // 16), ::max(1, n // 16)])
    for name, n, rank, k, over, seed in (("evd2_120_over0", 120, 40, 30, 0, 44), ("evd2_120_over5", 120, 40, 30, 5, 45),
                                         ("evd2_120_exact", 120, 20, 17, 5, 46)):
        B0 = orc.rand_low_rank(n, rank, rank, np.random.default_rng(seed))
        A = B0 @ B0.T
        A = 0.5 * (A + A.T)
        Vr, lr = revd.EVD2(raw.RS1(rsko.SkOpGA(), 1, rulaw.orth, 1))(A, k, np.nan, over, np.random.default_rng(7))
        Vo, lo = orc.EVD2(orc.RS1(orc.SkOpGA(), 1, orc.orth, 1))(A, k, np.nan, over, np.random.default_rng(7))
        ap = (Vr * lr) @ Vr.T
        print(f"{name:28s} EVD2 rank {lr.size}  d(approx) {relerr((Vo * lo) @ Vo.T, ap):.2e}  |A-VLV'|/|A| {relerr(ap, A):.2e}")
        assert lr.shape == lo.shape and relerr((Vo * lo) @ Vo.T, ap) < 1e-10
        np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="evd2", n=n, rank=rank, k=k, over=over, seed=seed,
                            spec=lr, approx_fro=np.linalg.norm(ap), err_fro=np.linalg.norm(A - ap), A_sha=digest(A),
                            approx_probe=ap[::max(1, n // 16), ::max(1, n // 16)])


def id_cases():
    """Interpolative / CUR decompositions (drivers/interpolative.py; test_osid.py, test_tsid.py, test_cur.py)."""
    import parla.drivers.interpolative as rid
    import parla.comps.interpolative as rci
    import parla.comps.sketchers.aware as raw
    for name, m, n, rank, k, over, npass, seed in (("id_tall_100x30", 100, 30, 30, 25, 4, 2, 51),
                                                   ("id_wide_30x100", 30, 100, 30, 27, 3, 2, 52),
                                                   ("id_exact_100x30", 100, 30, 5, 5, 1, 0, 53)):
        # (small shapes only: the fixture stores a hash of A, and the QR / GEMMs that build A are bit-reproducible
        #  across hosts with different BLAS thread counts only for small matrices)
        A = orc.rand_low_rank(m, n, rank, np.random.default_rng(seed))
        assert relerr(A, rmm.rand_low_rank(m, n, rank, np.random.default_rng(seed))) < 1e-13
        rs_r, rs_o = raw.RS1(rsko.SkOpGA(), npass, rulaw.orth, 1), orc.RS1(orc.SkOpGA(), npass, orc.orth, 1)
        fx = dict(m=m, n=n, rank=rank, k=k, over=over, num_pass=npass, seed=seed, A_sha=digest(A),
                  A_fro=np.linalg.norm(A))
        for axis in (0, 1):
            for tag, R_, O_ in (("osid1", rid.OSID1, orc.OSID1), ("osid2", rid.OSID2, orc.OSID2)):
                Mr, Pr = R_(rs_r)(A, k, over, axis, np.random.default_rng(7))
                Mo, Po = O_(rs_o)(A, k, over, axis, np.random.default_rng(7))
                assert np.array_equal(Pr, Po) and relerr(Mo, Mr) < 1e-8
                approx = Mr @ A[Pr, :] if axis == 0 else A[:, Pr] @ Mr
                fx[f"{tag}_ax{axis}_idx"] = Pr.astype(np.int32)
                fx[f"{tag}_ax{axis}_err"] = np.linalg.norm(A - approx)
            Pr = rci.ROCS1(rs_r)(A, k, over, axis, np.random.default_rng(7))
            assert np.array_equal(Pr, orc.ROCS1(rs_o)(A, k, over, axis, np.random.default_rng(7)))
            fx[f"rocs1_ax{axis}_idx"] = Pr.astype(np.int32)
        Z, Is, X, Js = rid.TSID1(rid.OSID1(rs_r))(A, k, over, np.random.default_rng(7))
        Zo, Iso, Xo, Jso = orc.TSID1(orc.OSID1(rs_o))(A, k, over, np.random.default_rng(7))
        assert np.array_equal(Is, Iso) and np.array_equal(Js, Jso) and relerr(Zo, Z) < 1e-8 and relerr(Xo, X) < 1e-8
        fx.update(tsid_Is=Is.astype(np.int32), tsid_Js=Js.astype(np.int32),
                  tsid_err=np.linalg.norm(A - Z @ A[Is, :][:, Js] @ X))
        Js, U, Is = rid.CUR1(rid.OSID1(rs_r))(A, k, over, np.random.default_rng(7))
        Jso, Uo, Iso = orc.CUR1(orc.OSID1(rs_o))(A, k, over, np.random.default_rng(7))
        assert np.array_equal(Is, Iso) and np.array_equal(Js, Jso)
        cur = A[:, Js] @ (U @ A[Is, :])
        assert relerr(A[:, Js] @ (Uo @ A[Is, :]), cur) < 1e-8
        fx.update(cur_Is=Is.astype(np.int32), cur_Js=Js.astype(np.int32), cur_err=np.linalg.norm(A - cur))
        print(f"{name:28s} OSID1 row/col err {fx['osid1_ax0_err'] / fx['A_fro']:.2e}/{fx['osid1_ax1_err'] / fx['A_fro']:.2e}  "
              f"TSID {fx['tsid_err'] / fx['A_fro']:.2e}  CUR {fx['cur_err'] / fx['A_fro']:.2e}")
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)


def philox_case():
    """Known-answer vectors for Philox4x32-10 (Random123 kat_vectors) + oracle stream samples."""
    from oracle import philox_ref as ph
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = ph.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want, (ctr, key, [hex(int(v)) for v in got])
    print("philox4x32-10                known-answer vectors OK")
    np.savez_compressed(os.path.join(OUT, "philox_kat.npz"),
                        ctr=np.array([c for c, _, _ in kat], dtype=np.uint32),
                        key=np.array([k for _, k, _ in kat], dtype=np.uint32),
                        out=np.array([o for _, _, o in kat], dtype=np.uint32))


if __name__ == "__main__":
    if "--only-id" in sys.argv:
        id_cases()
        sys.exit(0)
    if "--only-qb3" in sys.argv:
        qb3_evd2_cases()
        sys.exit(0)
    if "--only-srct" in sys.argv:
        srct_cases()
        sys.exit(0)
    if "--only-sps" in sys.argv:
        sps_cases()
        sys.exit(0)
    if "--only-big" in sys.argv:
        # BASELINE.json configs[1] at 1/16 of its rows (2^18 x 2048, d = 8192): ~1 min of reference + oracle time
        spo_case("spo_cfg2s_262144x2048", 262144, 2048, 0, 'sjlt', 'qr', 0.0, store_S=False, big=True)
        # configs[3] scaled: 2^14 x 2^11, k = 128, two power iterations, QB1 and QB2(blk = 32)
        lowrank_case("svd1_qb1_cfg4s_16384x2048", 16384, 2048, 1024, 128, 61, big=True, spectrum_param=50.0)
        lowrank_case("svd1_qb2_cfg4s_16384x2048", 16384, 2048, 1024, 128, 61, blk=32, tol=0.0, big=True,
                     spectrum_param=50.0)
        sys.exit(0)
    if "--only-rankdef" in sys.argv:
        spo_rankdef_case("spo_gauss_svd_rankdef_seed1", 1)
        spo_rankdef_case("spo_gauss_svd_rankdef_seed4", 4)
        sys.exit(0)
    if "--only-spu" in sys.argv:
        spu_case("spu1_sjlt_800x50", 800, 50, 31)
        spu_case("spu1_sjlt_2000x96", 2000, 96, 32, cond=1e5)
        sys.exit(0)
    philox_case()
    # sketch-and-precondition least squares: {SJLT, Gaussian} x {qr, svd, chol} x {delta}
    spo_case("spo_sjlt_qr_600x40", 600, 40, 11, 'sjlt', 'qr', 0.0)
    spo_case("spo_sjlt_svd_600x40", 600, 40, 11, 'sjlt', 'svd', 0.0)
    spo_case("spo_sjlt_chol_600x40", 600, 40, 11, 'sjlt', 'chol', 0.0)
    spo_case("spo_sjlt_qr_ridge_600x40", 600, 40, 12, 'sjlt', 'qr', 0.25)
    spo_case("spo_sjlt_svd_ridge_600x40", 600, 40, 12, 'sjlt', 'svd', 0.25)
    spo_case("spo_gauss_qr_500x37", 500, 37, 13, 'gauss', 'qr', 0.0)
    spo_case("spo_sjlt_qr_cond1e5_2000x64", 2000, 64, 14, 'sjlt', 'qr', 0.0, cond=1e5)
    spo_case("spo_sjlt_qr_odd_1531x77", 1531, 77, 15, 'sjlt', 'qr', 0.0, sf=3.3)
    # BASELINE.json configs[0]: 2^16 x 500, SJLT k=8, d=4n, tol 1e-12 (S regenerated + hash-checked)
    spo_case("spo_cfg1_65536x500", 65536, 500, 0, 'sjlt', 'qr', 0.0, store_S=False)
    # low-rank path
    lowrank_case("svd1_qb1_200x50", 200, 50, 15, 15, 21)
    lowrank_case("svd1_qb2_200x50", 200, 50, 45, 30, 22, blk=8, tol=0.0)
    lowrank_case("svd1_qb2_tol_200x50", 200, 50, 45, 50, 23, blk=5, tol=1e-3)
    lowrank_case("svd1_qb1_over_50x200", 50, 200, 15, 10, 24, over=5)
    lowrank_case("evd1_qb1_120", 120, 120, 12, 12, 25, evd=True)
    spu_case("spu1_sjlt_800x50", 800, 50, 31)
    spu_case("spu1_sjlt_2000x96", 2000, 96, 32, cond=1e5)
    sps_cases()
    srct_cases()
    qb3_evd2_cases()
    id_cases()
    spo_rankdef_case("spo_gauss_svd_rankdef_seed1", 1)
    spo_rankdef_case("spo_gauss_svd_rankdef_seed4", 4)
    print("golden fixtures written to", OUT, "(the large ones: rerun with --only-big)")
