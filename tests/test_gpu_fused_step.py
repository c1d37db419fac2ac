"""GPU: the fused vector phase of the LSQR iteration (pla_stream_pass_parts_f64 + pla_lsqr_fused_step_f64: reduce of
the pass's partials, t = M^T z, the recurrences of parla/comps/determiter/lsqr.py:421-526 and xw = M v_new in ONE
cluster launch) against the unfused chain of kernels it replaces, step by step and through whole solves."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
F64 = torch.float64

SHAPES = [(900, 6, 6), (3000, 30, 30), (5000, 77, 77), (2000, 64, 40), (65536, 500, 500), (4096, 1024, 1024),
          (3000, 1500, 600)]


def _problem(m, n, r, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(m, n, dtype=F64, device="cuda", generator=g)
    M = torch.randn(n, r, dtype=F64, device="cuda", generator=g) / np.sqrt(n)
    if n == r:
        M = torch.triu(M) + torch.eye(n, dtype=F64, device="cuda")       # like R^{-1}: upper triangular, dense storage
    b = torch.randn(m, dtype=F64, device="cuda", generator=g)
    return A, M, b


@pytest.mark.parametrize("m,n,r", SHAPES)
def test_one_fused_step_equals_reduce_precond_step_precond(m, n, r):
    from parla_b200 import kernels as K
    A, M, b = _problem(m, n, r, 7 * m + n)
    g = torch.Generator(device="cuda").manual_seed(1)
    # a running LSQR state: initialise from t0 = M^T A^T b, then compare ONE iteration done both ways
    zss0 = K.stream_pass(A, u=b.clone(), flags=K.PASS_AXPY)
    t0 = M.T @ zss0[:n]
    bsq = K.sumsq(b)
    state = {}
    for tag in ("ref", "fused"):
        x, v, w = (torch.empty(r, dtype=F64, device="cuda") for _ in range(3))
        ds = torch.zeros(K.LSQR_NDOUBLE, dtype=F64, device="cuda")
        is_ = torch.zeros(K.LSQR_NINT, dtype=torch.int32, device="cuda")
        hist = torch.full((10,), -1.0, dtype=F64, device="cuda")
        zs_init = torch.cat((t0, zss0[n:n + 1]))              # lsqr_init reads |u|^2 at index len(x)
        K.lsqr_init(t0, zs_init, bsq, 1e-14, 1e-14, 1e8, 10, None, x, v, w, ds, is_)
        state[tag] = dict(x=x, v=v, w=w, ds=ds, is_=is_, hist=hist, u=b.clone())
    assert int(state["ref"]["is_"][0]) == 0
    for it in range(3):
        # --- unfused: xw = M v ; pass + reduce ; t = M^T z ; step
        s = state["ref"]
        sc = s["ds"][K.LSQR_SA:K.LSQR_SA + 2]
        xw_ref = M @ s["v"]
        zss = K.stream_pass(A, w=xw_ref, u=s["u"], sc=sc, flags=K.PASS_DOT | K.PASS_AXPY)
        t_ref = M.T @ zss[:n]
        K.lsqr_step(t_ref, torch.cat((t_ref, zss[n:n + 1])), s["x"], s["v"], s["w"], s["ds"], s["is_"], s["hist"])
        # --- fused
        f = state["fused"]
        scf = f["ds"][K.LSQR_SA:K.LSQR_SA + 2]
        xw = (M @ f["v"]) if it == 0 else f["xw"]
        ws, nparts, ss_off = K.stream_pass_parts(A, w=xw, u=f["u"], sc=scf)
        assert nparts > 0 and ss_off >= nparts * n
        zss_f = torch.full((n + 1,), float("nan"), dtype=F64, device="cuda")
        t_f = torch.empty(r, dtype=F64, device="cuda")
        xw_new = torch.empty(n, dtype=F64, device="cuda")
        K.lsqr_fused_step(M, ws, nparts, ss_off, zss_f, t_f, f["x"], f["v"], f["w"], xw_new, f["ds"], f["is_"], f["hist"])
        f["xw"] = xw_new
        torch.cuda.synchronize()
        if it == 0:
            assert torch.equal(zss_f, zss), "the fused reduce must reproduce the reduce kernel's summation order"
        scale = float(torch.linalg.vector_norm(t_ref))
        assert float(torch.linalg.vector_norm(t_f - t_ref)) <= 1e-12 * scale
        for key in ("x", "v", "w"):
            ref, got = s[key], f[key]
            assert float(torch.linalg.vector_norm(got - ref)) <= 1e-11 * max(float(torch.linalg.vector_norm(ref)), 1e-300), key
        assert torch.allclose(f["ds"], s["ds"], rtol=1e-10, atol=1e-12 * float(s["ds"].abs().max()))
        assert torch.equal(f["is_"], s["is_"])
        want = M @ f["v"]
        assert float(torch.linalg.vector_norm(xw_new - want)) <= 1e-13 * max(float(torch.linalg.vector_norm(want)), 1e-300)
    # the zss-input form (nparts = 0) gives the same step as the partials form
    f = state["fused"]
    ws, nparts, ss_off = K.stream_pass_parts(A, w=f["xw"], u=f["u"], sc=f["ds"][K.LSQR_SA:K.LSQR_SA + 2])
    a = {k: f[k].clone() for k in ("x", "v", "w", "ds", "is_", "hist")}
    zz = torch.empty(n + 1, dtype=F64, device="cuda")
    tt, xa, xb = torch.empty(r, dtype=F64, device="cuda"), torch.empty(n, dtype=F64, device="cuda"), torch.empty(n, dtype=F64, device="cuda")
    K.lsqr_fused_step(M, ws, nparts, ss_off, zz, tt, a["x"], a["v"], a["w"], xa, a["ds"], a["is_"], a["hist"])
    K.lsqr_fused_step(M, None, 0, 0, zz.clone(), tt, f["x"], f["v"], f["w"], xb, f["ds"], f["is_"], f["hist"])
    assert torch.equal(a["x"], f["x"]) and torch.equal(a["v"], f["v"]) and torch.equal(xa, xb) and torch.equal(a["ds"], f["ds"])


def test_fused_step_is_a_noop_after_lsqr_has_stopped():
    from parla_b200 import kernels as K
    A, M, b = _problem(3000, 30, 30, 5)
    x, v, w = (torch.ones(30, dtype=F64, device="cuda") for _ in range(3))
    ds = torch.ones(K.LSQR_NDOUBLE, dtype=F64, device="cuda")
    is_ = torch.zeros(K.LSQR_NINT, dtype=torch.int32, device="cuda")
    is_[0] = 2
    hist = torch.zeros(4, dtype=F64, device="cuda")
    zss = torch.ones(31, dtype=F64, device="cuda")
    t, xw = torch.zeros(30, dtype=F64, device="cuda"), torch.full((30,), 7.0, dtype=F64, device="cuda")
    K.lsqr_fused_step(M, None, 0, 0, zss, t, x, v, w, xw, ds, is_, hist)
    torch.cuda.synchronize()
    assert bool((x == 1).all()) and bool((v == 1).all()) and bool((w == 1).all()) and bool((xw == 7).all())
    assert bool((ds == 1).all()) and int(is_[0]) == 2


@pytest.mark.parametrize("mode", ["qr", "svd"])
@pytest.mark.parametrize("m,n", [(20000, 120), (65536, 500), (3000, 33), (30000, 1000)])
def test_whole_solve_with_the_fused_vector_phase(m, n, mode):
    """SPO with the fused path switched on against the unfused default: same solution (1e-10), same iteration count
    (+-1), same error history (1e-6), and both equal to the dense least-squares solution."""
    import parla_b200 as rla
    from parla_b200.comps.determiter import lsqr as lsqr_mod
    rng = np.random.default_rng(m + n)
    A = rng.standard_normal((m, n)) * np.logspace(0, 2, n)
    b = A @ rng.standard_normal(n) + 0.1 * rng.standard_normal(m)
    Ad, bd = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    alg = rla.SPO(rla.SkOpSJ(8), 4, mode)
    old = lsqr_mod.USE_FUSED
    try:
        lsqr_mod.USE_FUSED = False
        c0 = rla.kernels.launch_count()
        x0, log0 = alg(Ad, bd, 0.0, 1e-12, 100, 11)
        c1 = rla.kernels.launch_count()
        lsqr_mod.USE_FUSED = True
        x1, log1 = alg(Ad, bd, 0.0, 1e-12, 100, 11)
        c2 = rla.kernels.launch_count()
    finally:
        lsqr_mod.USE_FUSED = old
    assert float(torch.linalg.vector_norm(x1 - x0) / torch.linalg.vector_norm(x0)) < 1e-10
    assert abs(log1.errors.size - log0.errors.size) <= 1
    k = min(log1.errors.size, log0.errors.size) - 2
    assert np.allclose(log1.errors[:k], log0.errors[:k], rtol=1e-6)
    assert log1.passes_over_A == log0.passes_over_A or abs(log1.passes_over_A - log0.passes_over_A) <= 1
    x_opt = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.linalg.norm(x1.cpu().numpy() - x_opt) <= 1e-9 * np.linalg.norm(x_opt)
    # two launches per iteration instead of seven or eight: at least four fewer per iteration
    iters = log1.errors.size - 1
    assert (c2 - c1) <= (c1 - c0) - 4 * (iters - 2), (c1 - c0, c2 - c1, iters)
