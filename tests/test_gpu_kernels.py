"""GPU: every C-ABI kernel against numpy on seeded inputs (bit-exact for index work, 1e-12-level
for fp64 reductions whose summation order differs from numpy's)."""
import os

import numpy as np
import pytest
import torch

from oracle import philox_ref as ph

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def K():
    from parla_b200 import kernels
    assert torch.cuda.is_available()
    return kernels


# ------------------------------------------------------------------ K4 streaming pass
@pytest.mark.parametrize("m,n", [(1, 1), (7, 3), (1000, 40), (4097, 500), (513, 77), (3000, 2048), (2049, 1025),
                                 (300, 4096), (150, 8192), (65, 2050), (10000, 64)])
def test_stream_pass_fused(K, m, n):
    rng = np.random.default_rng(m * 31 + n)
    A, w, u0 = rng.standard_normal((m, n)), rng.standard_normal(n), rng.standard_normal(m)
    sa, su = 0.7, -1.3
    u = dev(u0)
    zss = K.stream_pass(dev(A), w=dev(w), u=u, sa=sa, su=su, flags=K.PASS_DOT | K.PASS_AXPY).cpu().numpy()
    u_ref = sa * (A @ w) + su * u0
    assert rel(u, u_ref) < 1e-13
    assert np.linalg.norm(zss[:n] - A.T @ u_ref) <= 1e-12 * np.linalg.norm(A.T @ u_ref) + 1e-300
    assert abs(zss[n] - u_ref @ u_ref) <= 1e-13 * (u_ref @ u_ref)


def test_stream_pass_modes_and_device_scalars(K):
    rng = np.random.default_rng(5)
    m, n = 2500, 300
    A, x, b, g = rng.standard_normal((m, n)), rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(m)
    Ad = dev(A)
    y, zss = K.matvec(Ad, dev(x))                                     # DOT only
    assert rel(y, A @ x) < 1e-13 and abs(float(zss[n]) - (A @ x) @ (A @ x)) < 1e-13 * ((A @ x) @ (A @ x))
    assert np.all(zss[:n].cpu().numpy() == 0)
    zss = K.rmatvec(Ad, dev(b)).cpu().numpy()                          # AXPY only
    assert np.linalg.norm(zss[:n] - A.T @ b) < 1e-12 * np.linalg.norm(A.T @ b) and abs(zss[n] - b @ b) < 1e-13 * (b @ b)
    sc = dev(np.array([-1.0, 1.0]))                                   # residual + A^T g, scalars on device
    r = dev(b)
    zss = K.stream_pass(Ad, w=dev(x), u=r, g=dev(g), sc=sc,
                        flags=K.PASS_DOT | K.PASS_AXPY | K.PASS_AXPY_G).cpu().numpy()
    assert rel(r, b - A @ x) < 1e-13
    assert np.linalg.norm(zss[:n] - A.T @ g) < 1e-12 * np.linalg.norm(A.T @ g)
    y_nan = torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")   # su == 0: output only
    K.stream_pass(Ad, w=dev(x), u=y_nan, sa=2.0, su=0.0, flags=K.PASS_DOT)
    assert rel(y_nan, 2 * (A @ x)) < 1e-13
    stop = torch.ones(1, dtype=torch.int32, device="cuda")            # istop set => no-op
    r2 = dev(b)
    K.stream_pass(Ad, w=dev(x), u=r2, sa=5.0, su=0.0, flags=K.PASS_DOT, istop=stop)
    assert torch.equal(r2.cpu(), torch.from_numpy(b))


@pytest.mark.parametrize("m,n", [(65536, 500), (40000, 300), (30000, 1000), (50000, 77), (20000, 2048), (9000, 2049),
                                 (30000, 1700), (60000, 256)])
def test_stream_pass_single_product_modes_on_tall_matrices(K, m, n):
    """A^T u alone (adjoint_pass, rmatvec: no row sums, so nothing holds the warps of a consumer group together) and
    A w alone, with many ring wraps per CTA.  65536 x 500 (4 consumer groups, 6 stages) used to fail: a group met a
    stage only on every second fill and took an older completion for its own (csrc/stream_pass.cu: stage tags).  1700 columns: 2 groups on a 7-stage ring."""
    g = torch.Generator(device="cuda").manual_seed(m + n)
    A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    u = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    w = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    for _ in range(3):
        zss = K.stream_pass(A, u=u, flags=K.PASS_AXPY)
        ref = A.T @ u
        assert float(torch.linalg.vector_norm(zss[:n] - ref)) <= 1e-12 * float(torch.linalg.vector_norm(ref))
        assert abs(float(zss[n]) - float(u @ u)) <= 1e-12 * float(u @ u)
        y, zs2 = K.matvec(A, w)
        ref = A @ w
        assert float(torch.linalg.vector_norm(y - ref)) <= 1e-12 * float(torch.linalg.vector_norm(ref))
        u2 = u.clone()                                    # both products in one read (every ring shape with tags)
        zss = K.stream_pass(A, w=w, u=u2, sa=0.5, su=-1.0, flags=K.PASS_DOT | K.PASS_AXPY)
        uref = 0.5 * ref - u
        zref = A.T @ uref
        assert float(torch.linalg.vector_norm(u2 - uref)) <= 1e-12 * float(torch.linalg.vector_norm(uref))
        assert float(torch.linalg.vector_norm(zss[:n] - zref)) <= 1e-12 * float(torch.linalg.vector_norm(zref))


def test_stream_pass_strided_and_unaligned(K):
    rng = np.random.default_rng(6)
    big = rng.standard_normal((400, 130))
    Bd = dev(big)
    view = Bd[:, 3:103]                                               # lda = 130, base not 16B aligned
    w, u0 = rng.standard_normal(100), rng.standard_normal(400)
    u = dev(u0)
    zss = K.stream_pass(view, w=dev(w), u=u, sa=1.0, su=1.0, flags=3).cpu().numpy()
    Aref = big[:, 3:103]
    u_ref = Aref @ w + u0
    assert rel(u, u_ref) < 1e-13 and np.linalg.norm(zss[:100] - Aref.T @ u_ref) < 1e-12 * np.linalg.norm(Aref.T @ u_ref)


@pytest.mark.parametrize("m,n", [(300, 9000), (257, 16384), (200, 5001), (120, 12289)])
def test_stream_pass_wide_matrices(K, m, n):
    """n > 8192 (and odd n > 4096): column-blocked path, every flag combination the drivers use."""
    rng = np.random.default_rng(n)
    A, w, u0, gv = rng.standard_normal((m, n)), rng.standard_normal(n), rng.standard_normal(m), rng.standard_normal(m)
    Ad = dev(A)
    u = dev(u0)
    zss = K.stream_pass(Ad, w=dev(w), u=u, sa=0.7, su=-1.3, flags=K.PASS_DOT | K.PASS_AXPY).cpu().numpy()
    u_ref = 0.7 * (A @ w) + -1.3 * u0
    assert rel(u, u_ref) < 1e-13
    assert np.linalg.norm(zss[:n] - A.T @ u_ref) <= 1e-12 * np.linalg.norm(A.T @ u_ref)
    assert abs(zss[n] - u_ref @ u_ref) <= 1e-13 * (u_ref @ u_ref)
    y, zss = K.matvec(Ad, dev(w))
    assert rel(y, A @ w) < 1e-13 and abs(float(zss[n]) - (A @ w) @ (A @ w)) < 1e-13 * ((A @ w) @ (A @ w))
    zss = K.rmatvec(Ad, dev(u0)).cpu().numpy()
    assert np.linalg.norm(zss[:n] - A.T @ u0) < 1e-12 * np.linalg.norm(A.T @ u0) and abs(zss[n] - u0 @ u0) < 1e-13 * (u0 @ u0)
    sc = dev(np.array([-1.0, 1.0]))
    r = dev(u0)
    zss = K.stream_pass(Ad, w=dev(w), u=r, g=dev(gv), sc=sc, flags=K.PASS_DOT | K.PASS_AXPY | K.PASS_AXPY_G).cpu().numpy()
    assert rel(r, u0 - A @ w) < 1e-13 and np.linalg.norm(zss[:n] - A.T @ gv) < 1e-12 * np.linalg.norm(A.T @ gv)


def test_stream_pass_column_block_view_uses_row_copies(K):
    """A 16-byte aligned column block of a wider matrix (lda > n): one bulk copy per row of a tile."""
    rng = np.random.default_rng(61)
    big = rng.standard_normal((3000, 1536))
    view = dev(big)[:, 512:1024]
    w, u0 = rng.standard_normal(512), rng.standard_normal(3000)
    u = dev(u0)
    zss = K.stream_pass(view, w=dev(w), u=u, sa=1.0, su=1.0, flags=3).cpu().numpy()
    Aref = big[:, 512:1024]
    u_ref = Aref @ w + u0
    assert rel(u, u_ref) < 1e-13 and np.linalg.norm(zss[:512] - Aref.T @ u_ref) < 1e-12 * np.linalg.norm(Aref.T @ u_ref)


def test_spo_wide_matrix_end_to_end(K):
    """SPO on n > 8192 columns (the round-1 hard limit of the streaming pass): small m keeps it quick."""
    import parla_b200 as rla
    g = torch.Generator(device="cuda").manual_seed(5)
    m, n = 12000, 8448
    A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    x0 = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    b = A @ x0
    x, log = rla.SPO(rla.SkOpSJ(8), 1.2, 'qr')(A, b, 0.0, 1e-12, 60, 3)
    assert float(torch.linalg.vector_norm(x - x0) / torch.linalg.vector_norm(x0)) < 1e-8


def test_stream_pass_deterministic(K):
    rng = np.random.default_rng(7)
    A, w = dev(rng.standard_normal((20000, 512))), dev(rng.standard_normal(512))
    outs = []
    for _ in range(3):
        u = torch.zeros(20000, dtype=torch.float64, device="cuda")
        outs.append(K.stream_pass(A, w=w, u=u, sa=1.0, su=0.0, flags=3).clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


# ------------------------------------------------------------------ K3b trsv
@pytest.mark.parametrize("n", [1, 5, 32, 33, 100, 500, 2048])
@pytest.mark.parametrize("trans", [False, True])
def test_trsv(K, n, trans):
    rng = np.random.default_rng(n)
    R = np.linalg.qr(rng.standard_normal((2 * n + 3, n)))[1]              # well conditioned (random triangular is not)
    R_full = R + np.tril(rng.standard_normal((n, n)), -1)              # junk below the diagonal is ignored
    b = rng.standard_normal(n)
    x = K.trsv_upper(dev(R_full), dev(b), trans=trans)
    x_ref = np.linalg.solve(R.T if trans else R, b)
    assert rel(x, x_ref) < 1e-11
    # strided R (leading dimension n+1) and in-place
    W = np.zeros((n, n + 1)); W[:, :n] = R_full
    Wd = dev(W)
    bd = dev(b)
    K.trsv_upper(Wd[:, :n], bd, trans=trans, out=bd)
    assert rel(bd, x_ref) < 1e-11


@pytest.mark.parametrize("n", [1, 31, 32, 33, 100, 257, 1024, 2048])
def test_trtri_upper(K, n):
    rng = np.random.default_rng(n + 5)
    R = np.linalg.qr(rng.standard_normal((2 * n + 3, n)))[1]
    W = np.zeros((n, n + 1)); W[:, :n] = R + np.tril(rng.standard_normal((n, n)), -1)   # strided, junk below diag
    X = K.trtri_upper(dev(W)[:, :n]).cpu().numpy()
    assert np.allclose(X, np.triu(X))
    assert np.linalg.norm(X @ R - np.eye(n)) < 1e-11 * n


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K_", [(1, 1, 1), (37, 29, 53), (128, 128, 64), (200, 130, 1000), (64, 48, 20000),
                                    (257, 2, 130)])
def test_gemm(K, ta, tb, M, N, K_):
    rng = np.random.default_rng(M + 7 * N + 13 * K_ + ta + 2 * tb)
    A = rng.standard_normal((K_, M) if ta else (M, K_))
    B = rng.standard_normal((N, K_) if tb else (K_, N))
    C0 = rng.standard_normal((M, N))
    ref = 0.5 * (A.T if ta else A) @ (B.T if tb else B) + 2.0 * C0
    C = dev(C0)
    K.gemm(dev(A), dev(B), transa=bool(ta), transb=bool(tb), alpha=0.5, beta=2.0, out=C)
    assert rel(C, ref) < 1e-13
    C2 = K.gemm(dev(A), dev(B), transa=bool(ta), transb=bool(tb))
    assert rel(C2, (A.T if ta else A) @ (B.T if tb else B)) < 1e-13


def test_gemm_strided_views(K):
    rng = np.random.default_rng(3)
    big = dev(rng.standard_normal((300, 90)))
    A = big[:, 1:60]                                                   # lda 90, unaligned
    B = dev(rng.standard_normal((59, 33)))
    out = torch.zeros(300, 40, dtype=torch.float64, device="cuda")
    K.gemm(A, B, out=out[:, 2:35])
    assert rel(out[:, 2:35], A.cpu().numpy() @ B.cpu().numpy()) < 1e-13
    assert float(out[:, :2].abs().sum()) == 0 and float(out[:, 35:].abs().sum()) == 0


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K_", [(131, 257, 77), (128, 1921, 301), (515, 129, 128)])
def test_gemm_odd_extents_on_an_even_pitch(K, ta, tb, M, N, K_):
    """Odd extents in buffers whose rows are 16-byte aligned (the d x (n + 1) sketch inside the QR) take the
    16-byte copy path with a half-filled last slot; the pad column holds NaN, so touching it poisons the result."""
    rng = np.random.default_rng(M + N + K_ + ta + 2 * tb)

    def padded(x):
        r, c = x.shape
        buf = torch.full((r, c + (c & 1)), float("nan"), dtype=torch.float64, device="cuda")
        buf[:, :c] = dev(x)
        return buf[:, :c]

    A = rng.standard_normal((K_, M) if ta else (M, K_))
    B = rng.standard_normal((N, K_) if tb else (K_, N))
    C0 = rng.standard_normal((M, N))
    ref = 0.5 * (A.T if ta else A) @ (B.T if tb else B) - C0
    C = padded(C0)
    K.gemm(padded(A), padded(B), transa=bool(ta), transb=bool(tb), alpha=0.5, beta=-1.0, out=C)
    assert rel(C, ref) < 1e-13


# ------------------------------------------------------------------ Philox / Gaussian sketch
def test_philox_fill_matches_oracle_definition(K):
    S = K.philox_normal_fill(37, 1001, seed=2024, scale=0.5, row_offset=3, col_offset=6).cpu().numpy()
    ref = ph.gaussian_block(2024, 3, 37, 6, 1001, scale=0.5)
    assert np.max(np.abs(S - ref)) < 5e-6           # fp32 fast-math Box-Muller vs float64 libm
    big = K.philox_normal_fill(256, 8192, seed=11).cpu().numpy()
    assert abs(big.mean()) < 3e-3 and abs(big.std() - 1) < 3e-3
    from scipy import stats
    assert stats.kstest(big.reshape(-1)[::7], "norm").pvalue > 1e-3


@pytest.mark.parametrize("m,n,d,with_b", [(1000, 40, 160, True), (5003, 129, 300, True), (4096, 256, 512, False),
                                          (70000, 64, 128, True), (6000, 128, 300, True), (4100, 256, 512, True)])
def test_sketch_gauss_equals_materialised_operator(K, m, n, d, with_b):
    rng = np.random.default_rng(m + n)
    A, b = rng.standard_normal((m, n)), rng.standard_normal(m)
    scale = 1 / np.sqrt(d)
    S = K.philox_normal_fill(d, m, seed=77, scale=scale).cpu().numpy()
    out = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")
    K.sketch_gauss(dev(A), d, 77, scale, out, bvec=dev(b) if with_b else None)
    assert rel(out[:, :n], S @ A) < 1e-13
    if with_b:
        assert rel(out[:, n], S @ b) < 1e-13
    # column offset = row shard: two half-sketches add up to the full one
    h = (m // 8) * 4
    o1 = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")
    o2 = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")
    K.sketch_gauss(dev(A[:h]), d, 77, scale, o1, col_offset=0)
    K.sketch_gauss(dev(A[h:]), d, 77, scale, o2, col_offset=h)
    assert rel(o1[:, :n] + o2[:, :n], S @ A) < 1e-13


# ------------------------------------------------------------------ SJLT
@pytest.mark.parametrize("m,n,d,k", [(500, 40, 160, 8), (3001, 77, 254, 8), (20000, 512, 2048, 8), (64, 9, 5, 8),
                                     (1000, 300, 1200, 1)])
def test_sjlt_apply_matches_scipy(K, m, n, d, k):
    import scipy.sparse as sps
    rng = np.random.default_rng(m + d)
    kk = min(k, d) if d >= k else k
    rows = np.stack([rng.choice(d, kk, replace=d < kk) for _ in range(m)]).astype(np.int32)
    signs = rng.choice([-1, 1], size=(m, kk)).astype(np.int8)
    S = sps.coo_matrix((signs.reshape(-1) / np.sqrt(kk), (rows.reshape(-1), np.repeat(np.arange(m), kk))),
                       shape=(d, m)).tocsc()
    A, b = rng.standard_normal((m, n)), rng.standard_normal(m)
    plan = K.SjltPlan(dev(rows), dev(signs), d, validate=True)
    out = torch.full((d, n + 1), np.nan, dtype=torch.float64, device="cuda")
    plan.apply(dev(A), 1 / np.sqrt(kk), out, bvec=dev(b), out_b=out[:, n])
    assert rel(out[:, :n], S @ A) < 1e-13 and rel(out[:, n], S @ b) < 1e-13
    plan.apply(dev(A), 1 / np.sqrt(kk), out, bvec=dev(b), out_b=out[:, n], accumulate=True)
    assert rel(out[:, :n], 2 * (S @ A)) < 1e-13
    out2 = torch.empty(d, n + 1, dtype=torch.float64, device="cuda")       # deterministic
    plan2 = K.SjltPlan(dev(rows), dev(signs), d)
    plan2.apply(dev(A), 1 / np.sqrt(kk), out2, bvec=dev(b), out_b=out2[:, n])
    plan2.apply(dev(A), 2 / np.sqrt(kk), out, bvec=dev(b), out_b=out[:, n])
    assert torch.equal(out, 2 * out2)


def test_sjlt_oversized_bucket_is_sorted_not_skipped(K):
    """A (contrived) operator whose nonzeros pile up in one destination row: the bucket is longer than the shared-
    memory sort buffer (8192) and is sorted in global memory -- same result on every run, bit for bit."""
    import scipy.sparse as sps
    rng = np.random.default_rng(77)
    m, n, d = 20000, 33, 16
    rows = np.full((m, 1), 3, dtype=np.int32)
    rows[::7, 0] = 11
    signs = rng.choice([-1, 1], size=(m, 1)).astype(np.int8)
    S = sps.coo_matrix((signs.reshape(-1).astype(float), (rows.reshape(-1), np.arange(m))), shape=(d, m)).tocsc()
    A = rng.standard_normal((m, n))
    outs = []
    for _ in range(3):
        plan = K.SjltPlan(dev(rows), dev(signs), d, validate=True)
        out = torch.empty(d, n, dtype=torch.float64, device="cuda")
        plan.apply(dev(A), 1.0, out)
        outs.append(out)
    assert rel(outs[0], S @ A) < 1e-13
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


def test_sjlt_plan_rejects_bad_indices(K):
    rows = dev(np.array([[0, 7], [1, 2]], dtype=np.int32))
    signs = dev(np.ones((2, 2), dtype=np.int8))
    with pytest.raises(ValueError):
        K.SjltPlan(rows, signs, 5, validate=True)


def test_sjlt_generate_bit_exact(K):
    for d, m, k, seed, off in [(200, 3000, 8, 5, 0), (17, 500, 8, 6, 1000), (4, 50, 8, 7, 0), (8192, 2000, 8, 8, 2 ** 33)]:
        rows, signs = K.sjlt_generate(d, m, k, seed, col_offset=off)
        r_ref, s_ref = ph.sjlt_columns(d, m, k, seed, col_offset=off)
        assert np.array_equal(rows.cpu().numpy(), r_ref) and np.array_equal(signs.cpu().numpy(), s_ref)


# ------------------------------------------------------------------ Householder QR
@pytest.mark.parametrize("M,N", [(1, 1), (5, 3), (40, 40), (100, 17), (300, 64), (1000, 130), (5000, 33), (257, 200)])
def test_qr_economic(K, M, N):
    rng = np.random.default_rng(M * 3 + N)
    Y = rng.standard_normal((M, N))
    Q, R = K.qr_economic(dev(Y))
    Q, R = Q.cpu().numpy(), R.cpu().numpy()
    k = min(M, N)
    assert Q.shape == (M, k) and R.shape == (k, N)
    assert np.linalg.norm(Q.T @ Q - np.eye(k)) < 1e-13 * k
    assert np.linalg.norm(Q @ R - Y) < 1e-13 * np.linalg.norm(Y)
    assert np.allclose(R, np.triu(R))
    # LAPACK conventions: same R (incl. signs) as scipy's dgeqrf
    import scipy.linalg as sla
    R_ref = sla.qr(Y, mode='economic')[1]
    assert np.linalg.norm(R - R_ref) < 1e-11 * np.linalg.norm(R_ref)


def test_qr_rank_deficient_and_rhs_column(K):
    rng = np.random.default_rng(9)
    Y = rng.standard_normal((400, 6)) @ rng.standard_normal((6, 20))      # rank 6
    Q, R = K.qr_economic(dev(Y))
    Q = Q.cpu().numpy()
    assert np.linalg.norm(Q.T @ Q - np.eye(20)) < 1e-12
    assert np.linalg.norm(Q @ R.cpu().numpy() - Y) < 1e-12 * np.linalg.norm(Y)
    # [A | b]: factor the first n columns, get Q^T b in the last one
    A, b = rng.standard_normal((700, 50)), rng.standard_normal(700)
    W = dev(np.column_stack([A, b]))
    K.geqrf(W, 50)
    import scipy.linalg as sla
    Qr, Rr = sla.qr(A, mode='economic')
    assert rel(torch.triu(W[:50, :50]), Rr) < 1e-12
    assert rel(W[:50, 50], Qr.T @ b) < 1e-12
    x = K.trsv_upper(W[:50, :50], W[:50, 50].contiguous())
    assert rel(x, np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-11


def test_sumsq(K):
    rng = np.random.default_rng(1)
    for n in (1, 1000, 1 << 20):
        x = rng.standard_normal(n)
        assert abs(float(K.sumsq(dev(x))) - x @ x) <= 1e-13 * (x @ x)
