"""GPU: the sketch-and-precondition driver against (1) the golden fixtures (reference outputs) with
the reference's own sketching operator replayed, (2) the oracle run live on the same inputs, and
(3) the mathematical properties the reference's own test-suite asserts
(parla/tests/test_drivers/test_optim/test_overdet_least_squares.py:142-251).

Stated fp64 tolerances (BASELINE.json north_star): 1e-10 relative on x and on |Ax - b|."""
import warnings

import numpy as np
import pytest
import torch

from oracle import parla_oracle as orc
from tests.helpers import (SPO_FIXTURES, SPO_RANKDEF_FIXTURES, SPU_FIXTURES, Replay, load_golden,
                           problem_from_fixture, sjlt_from_fixture, spu_problem_from_fixture)

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
TOL_X = 1e-10


@pytest.fixture(scope="module")
def rla():
    import parla_b200
    return parla_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def replayed_operator(fx, d, m):
    if str(fx["sketch"]) == "sjlt":
        return sjlt_from_fixture(fx, d, m)
    return orc.gaussian_operator(d, m, np.random.default_rng(int(fx["rng_seed"])))


@pytest.mark.parametrize("name", SPO_FIXTURES)
def test_spo_matches_reference_fixture(rla, name):
    fx = load_golden(name)
    A, b = problem_from_fixture(fx)
    m, n = A.shape
    d = int(float(fx["sf"]) * n)
    S = replayed_operator(fx, d, m)
    alg = rla.SPO(Replay(S), float(fx["sf"]), str(fx["mode"]))
    x, log = alg(dev(A), dev(b), float(fx["delta"]), float(fx["tol"]), int(fx["iter_lim"]), None)
    x = x.cpu().numpy()
    assert np.linalg.norm(x - fx["x"]) <= TOL_X * np.linalg.norm(fx["x"])
    r = np.linalg.norm(A @ x - b)
    assert abs(r - float(fx["resid_norm"])) <= TOL_X * float(fx["resid_norm"])
    assert abs((log.errors.size - 1) - (fx["errors"].size - 1)) <= 1           # iteration count +-1
    k = min(log.errors.size, fx["errors"].size)
    assert np.allclose(log.errors[:k], fx["errors"][:k], rtol=1e-6, atol=1e-11 * fx["errors"][0])
    assert log.times.size == log.errors.size and log.time_sketch > 0 and log.time_iterate > 0


def test_spo_cfg1_full_size_parity(rla):
    """BASELINE.json configs[0] (2^16 x 500, SJLT k=8, d=4n, tol 1e-12) against the reference's output."""
    fx = load_golden("spo_cfg1_65536x500")
    A, b = problem_from_fixture(fx)
    S = sjlt_from_fixture(fx, 2000, 65536)
    x, log = rla.SPO(Replay(S), 4, 'qr')(dev(A), dev(b), 0.0, 1e-12, 100, None)
    x = x.cpu().numpy()
    assert np.linalg.norm(x - fx["x"]) <= TOL_X * np.linalg.norm(fx["x"])
    assert abs(np.linalg.norm(A @ x - b) - float(fx["resid_norm"])) <= TOL_X * float(fx["resid_norm"])
    assert abs(log.errors.size - fx["errors"].size) <= 1
    assert log.passes_over_A <= log.iters + 4            # sketch + presolve/init + iterations + final


def test_spo_cfg2_scaled_parity(rla):
    """BASELINE.json configs[1] at 1/16 of its rows -- 2^18 x 2048, SJLT k = 8, d = 4n = 8192, tol 1e-12 -- against
    the REFERENCE's output (oracle/make_golden.py --only-big; the reference's own SJLT is regenerated from its
    seed and hash-checked, A is hash-checked).  This is the size at which the windowed SJLT apply, the 128-column
    block-reflector QR and the split-K DMMA GEMMs all engage."""
    fx = load_golden("spo_cfg2s_262144x2048")
    A, b = problem_from_fixture(fx)
    m, n = A.shape
    S = sjlt_from_fixture(fx, 4 * n, m)
    x, log = rla.SAP1(Replay(S), 4)(dev(A), dev(b), 0.0, 1e-12, 100, None)
    x = x.cpu().numpy()
    assert np.linalg.norm(x - fx["x"]) <= TOL_X * np.linalg.norm(fx["x"])
    r = np.linalg.norm(A @ x - b)
    assert abs(r - float(fx["resid_norm"])) <= TOL_X * float(fx["resid_norm"])
    assert abs(log.errors.size - fx["errors"].size) <= 1                       # iteration count +-1
    k = min(log.errors.size, fx["errors"].size)
    assert np.allclose(log.errors[:k], fx["errors"][:k], rtol=1e-6, atol=1e-11 * fx["errors"][0])


@pytest.mark.parametrize("name", SPO_RANKDEF_FIXTURES)
def test_spo_svd_rank_deficient_matches_reference(rla, name):
    """The reference's `consistent_lowrank` problem (test_overdet_least_squares.py:23-34, :408-411 with the
    intended tol 1e-12): rank 5 of 10 columns, SPO(mode='svd') truncates the preconditioner to the numerical rank
    (preconditioning.py:74-77); with seed 4 the presolve is rejected because R is non-square
    (least_squares.py:348-351) and LSQR starts from the origin."""
    fx = load_golden(name)
    A, b, S = fx["A"], fx["b"], fx["S"]
    for alg in (rla.SPO(Replay(S), 3, 'svd'), rla.SAP2(Replay(S), 3)):
        x, log = alg(dev(A), dev(b), 0.0, float(fx["tol"]), int(fx["iter_lim"]), None)
        x = x.cpu().numpy()
        assert np.linalg.norm(x - fx["x"]) <= TOL_X * np.linalg.norm(fx["x"])
        assert np.linalg.norm(x - fx["x_minnorm"]) <= 1e-9 * np.linalg.norm(fx["x_minnorm"])   # minimum-norm solution
        assert np.linalg.norm(A @ x - b) <= 1e-12 * np.linalg.norm(b)
        # The presolve is accepted iff |A x_ske - b| <= 1e-15 |b| when R is non-square (least_squares.py:348): a
        # round-off-level threshold, so an equally accurate x_ske can land on either side of it.  Same branch as the
        # reference => same history; other branch => the presolve was accepted (one LSQR step, already converged).
        if log.errors.size == fx["errors"].size:
            big = fx["errors"] > 1e-9 * fx["errors"][0]
            assert np.allclose(log.errors[big], fx["errors"][big], rtol=1e-6)
        else:
            assert log.errors.size == 2 and log.errors[-1] <= 1e-12 * log.errors[0]
            assert abs(log.errors[0] - fx["errors"][0]) <= 1e-9 * fx["errors"][0]


def test_lsqr_iteration_as_cuda_graph_gives_identical_results(rla):
    """The optional CUDA-graph replay of the LSQR iteration (PLA_LSQR_GRAPH=1; off by default because it measured
    slower) must be bit-identical to the eager loop: same kernels, same order, same buffers."""
    from parla_b200.comps.determiter import lsqr as lsqr_mod
    rng = np.random.default_rng(17)
    A = rng.standard_normal((20000, 120)) * np.logspace(0, 3, 120)
    b = rng.standard_normal(20000)
    Ad, bd = dev(A), dev(b)
    for delta in (0.0, 0.3):
        x_eager, log_eager = rla.SAP1(rla.SkOpSJ(8), 4)(Ad, bd, delta, 1e-12, 100, 3)
        lsqr_mod.USE_GRAPH = True
        try:
            x_graph, log_graph = rla.SAP1(rla.SkOpSJ(8), 4)(Ad, bd, delta, 1e-12, 100, 3)
        finally:
            lsqr_mod.USE_GRAPH = False
        assert torch.equal(x_eager, x_graph) and np.array_equal(log_eager.errors, log_graph.errors)
        assert log_graph.passes_over_A == log_eager.passes_over_A


def test_sap1_sap2_are_spo_modes(rla):
    """CHANGELOG.md:57 names: SAP1 == SPO(mode='qr'), SAP2 == SPO(mode='svd'); bit-identical results."""
    rng = np.random.default_rng(21)
    A = rng.standard_normal((3000, 70)); b = rng.standard_normal(3000)
    Ad, bd = dev(A), dev(b)
    for cls, mode in ((rla.SAP1, 'qr'), (rla.SAP2, 'svd')):
        alg = cls(rla.SkOpSJ(8), 4)
        assert isinstance(alg, rla.SPO) and alg.mode == mode
        x1, l1 = alg(Ad, bd, 0.0, 1e-12, 100, 5)
        x2, l2 = rla.SPO(rla.SkOpSJ(8), 4, mode)(Ad, bd, 0.0, 1e-12, 100, 5)
        assert torch.equal(x1, x2) and np.array_equal(l1.errors, l2.errors)
        assert alg.exec.__func__ is rla.SPO.__call__


def test_sso1_rank_deficient_sketch_is_min_norm(rla):
    """SSO1 calls la.lstsq (gelsd: rank revealing, minimum norm; least_squares.py:184-186).  A sketch with
    dependent columns must give the pseudo-inverse solution, not inf/nan from a triangular solve."""
    rng = np.random.default_rng(6)
    A = rng.standard_normal((1500, 24))
    A[:, [7, 19]] = 0.0           # rank 22 with EXACTLY zero singular values (dependent columns would leave O(eps)
                                  # ones, which gelsd's cut-off eps * sigma_max keeps or drops by luck)
    b = rng.standard_normal(1500)
    S = orc.sjlt_operator(96, 1500, np.random.default_rng(2), 8)
    x_ref, _ = orc.SSO1(Replay(S), 4)(A, b, 0.0, np.nan, 1, None)
    x, _ = rla.SSO1(Replay(S), 4)(dev(A), dev(b), 0.0, np.nan, 1, None)
    x = x.cpu().numpy()
    assert np.all(np.isfinite(x)) and abs(x[7]) <= 1e-14 and abs(x[19]) <= 1e-14
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref)


@pytest.mark.parametrize("mode", ["qr", "svd", "chol"])
@pytest.mark.parametrize("delta", [0.0, 0.5])
def test_spo_against_live_oracle(rla, mode, delta):
    rng = np.random.default_rng(100)
    m, n = 3000, 120
    A = rng.standard_normal((m, n)) * np.logspace(0, 3, n)
    b = A @ rng.standard_normal(n) + rng.standard_normal(m)
    S = orc.sjlt_operator(4 * n, m, np.random.default_rng(3), 8)
    x_ref, log_ref = orc.SPO(Replay(S), 4, mode)(A, b, delta, 1e-12, 100, None)
    x, log = rla.SPO(Replay(S), 4, mode)(dev(A), dev(b), delta, 1e-12, 100, None)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= TOL_X * np.linalg.norm(x_ref)
    assert abs(log.errors.size - log_ref.errors.size) <= 1
    assert abs(log.errors[0] - log_ref.errors[0]) <= 1e-9 * log_ref.errors[0]


def test_spo_host_buffers_in_numpy_out(rla):
    rng = np.random.default_rng(8)
    A = rng.standard_normal((2000, 60)); b = rng.standard_normal(2000)
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(A, b, 0.0, 1e-12, 100, 3)
    assert isinstance(x, np.ndarray)
    x_opt = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.linalg.norm(x - x_opt) <= 1e-9 * np.linalg.norm(x_opt)


@pytest.mark.parametrize("gen_name", ["SkOpSJ", "SkOpGA"])
def test_spo_streamed_upload_of_pinned_host_buffers(rla, gen_name):
    """Host-resident (pinned) A, b: uploaded in row blocks on a copy stream, each block sketched as it arrives
    (the e2e path of bench.py).  Same x as with device-resident inputs, to the stated tolerance."""
    rng = np.random.default_rng(31)
    m, n = 70000, 96
    A = rng.standard_normal((m, n)) * np.logspace(0, 2, n)
    b = A @ rng.standard_normal(n) + 0.1 * rng.standard_normal(m)
    gen = getattr(rla, gen_name)
    gen = gen(8) if gen_name == "SkOpSJ" else gen()
    Ah, bh = torch.from_numpy(A).pin_memory(), torch.from_numpy(b).pin_memory()
    alg = rla.SPO(gen, 4, 'qr')
    xh, _ = alg(Ah, bh, 0.0, 1e-12, 100, 9)
    assert isinstance(xh, torch.Tensor) and not xh.is_cuda
    up = alg.last_upload
    assert up is not None and up["bytes"] == m * n * 8 + m * 8 and up["seconds"] > 0
    xd, _ = rla.SPO(gen, 4, 'qr')(dev(A), dev(b), 0.0, 1e-12, 100, 9)
    assert np.linalg.norm(xh.numpy() - xd.cpu().numpy()) <= TOL_X * np.linalg.norm(xd.cpu().numpy())
    x_opt = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.linalg.norm(xh.numpy() - x_opt) <= 1e-9 * np.linalg.norm(x_opt)
    xn, _ = alg(A, b, 0.0, 1e-12, 100, 9)                      # pageable numpy buffers take the same path
    assert isinstance(xn, np.ndarray) and np.linalg.norm(xn - xh.numpy()) <= TOL_X * np.linalg.norm(xn)


@pytest.mark.parametrize("gen_name", ["SkOpSJ", "SkOpGA", "sjlt_operator", "gaussian_operator"])
@pytest.mark.parametrize("mode", ["qr", "svd"])
def test_native_operators_convergence_rate(rla, gen_name, mode):
    """Reference property test (test_overdet_least_squares.py:234-251): log-linear convergence with
    R^2 >= 0.95, slope < -0.3, final error <= 1e-6 -- with the Philox-native operators."""
    rng = np.random.default_rng(897809809)
    m, n = 1000, 100
    spec = np.concatenate([1e5 + rng.random(30), 1 + rng.random(70)])
    U = orc.orthonormal_operator(m, n, rng)
    Vt = orc.orthonormal_operator(n, n, rng)
    A = (U * spec) @ Vt
    xt = np.concatenate([rng.standard_normal(30) / 1e5, rng.standard_normal(70)])
    b_orth = rng.standard_normal(m) * 1e2
    b_orth -= U @ (U.T @ b_orth)
    b = A @ xt + b_orth
    gen = getattr(rla, gen_name)
    gen = gen(8) if gen_name == "SkOpSJ" else (gen() if gen_name == "SkOpGA" else gen)
    x, log = rla.SPO(gen, 3, mode)(dev(A), dev(b), 0.0, 1e-12, 100, np.random.default_rng(34998751340))
    errs = log.errors[1:]
    t = np.arange(errs.size)
    slope, icpt = np.polyfit(t, np.log(errs), 1)
    fit = slope * t + icpt
    r2 = 1 - np.sum((np.log(errs) - fit) ** 2) / np.sum((np.log(errs) - np.log(errs).mean()) ** 2)
    assert r2 >= 0.95 and slope < -0.3 and log.errors[-1] <= 1e-6
    x = x.cpu().numpy()
    res = A @ x - b                                               # test_residual_proj, :171-180
    assert np.linalg.norm(U @ (U.T @ res)) / np.linalg.norm(res) <= 1e-6
    assert abs(np.linalg.norm(Vt @ x) - np.linalg.norm(Vt @ xt)) <= 1e-6 * (1 + np.linalg.norm(Vt @ xt))


def test_spo_edge_cases(rla):
    rng = np.random.default_rng(1)
    # consistent system: presolve is (nearly) exact, LSQR stops immediately or after a step
    A = rng.standard_normal((500, 20)); xt = rng.standard_normal(20)
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(dev(A), dev(A @ xt), 0.0, 1e-12, 50, 1)
    assert np.linalg.norm(x.cpu().numpy() - xt) <= 1e-10 * np.linalg.norm(xt)
    # b orthogonal to range(A): x = 0, zero vector beats the presolve (rel_err >= 1 branch)
    Q = np.linalg.qr(rng.standard_normal((500, 21)))[0]
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(dev(Q[:, :20]), dev(Q[:, 20]), 0.0, 1e-12, 50, 1)
    assert np.linalg.norm(x.cpu().numpy()) <= 1e-10
    # d clipped to m with a warning (dim_checks), square-ish problem
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        A = rng.standard_normal((60, 50)); b = rng.standard_normal(60)
        x, _ = rla.SPO(rla.SkOpGA(), 4, 'qr')(dev(A), dev(b), 0.0, 1e-12, 200, 1)
        assert any("embedding dimension" in str(i.message) for i in w)
    assert np.linalg.norm(x.cpu().numpy() - np.linalg.lstsq(A, b, rcond=None)[0]) <= 1e-8
    with pytest.raises(ValueError):
        rla.SPO(rla.SkOpSJ(8), 4, 'nope')(dev(A), dev(b), 0.0, 1e-12, 10, 1)
    # iteration limit reached: istop 7 after exactly iter_lim steps
    A = rng.standard_normal((2000, 100)) * np.logspace(0, 6, 100); b = rng.standard_normal(2000)
    x, log = rla.SPO(rla.SkOpSJ(8), 1.2, 'qr')(dev(A), dev(b), 0.0, 1e-15, 5, 1)
    assert log.errors.size - 1 == 5


def test_sso1_sketch_and_solve(rla):
    rng = np.random.default_rng(2)
    A = rng.standard_normal((4000, 50)); b = A @ rng.standard_normal(50) + 0.01 * rng.standard_normal(4000)
    S = orc.sjlt_operator(300, 4000, np.random.default_rng(5), 8)
    x_ref, _ = orc.SSO1(Replay(S), 6)(A, b, 0.0, np.nan, 1, None)
    x, log = rla.SSO1(Replay(S), 6)(dev(A), dev(b), 0.0, np.nan, 1, None)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-10 * np.linalg.norm(x_ref)
    assert set(log) == {"time_sketch", "time_solve"}
    x_ref, _ = orc.SSO1(Replay(S), 6)(A, b, 0.3, np.nan, 1, None)
    x, _ = rla.SSO1(Replay(S), 6)(dev(A), dev(b), 0.3, np.nan, 1, None)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-10 * np.linalg.norm(x_ref)


def test_spo_full_size_properties(rla):
    """BASELINE.json configs[1] shape (2^22 x 2048 fp64, d = 4n, SJLT) through size-independent
    properties: b = A x0 + noise  =>  the normal-equation residual is at round-off level relative to
    |A||r|, x is within the noise cone of x0, and the error history decays linearly."""
    free, _ = torch.cuda.mem_get_info()
    m, n = (1 << 22, 2048) if free > 100e9 else (1 << 19, 2048)
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=g)
    x0 = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    from parla_b200 import kernels as K
    b, _ = K.matvec(A, x0)
    b += 0.1 * torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(A, b, 0.0, 1e-12, 100, 3)
    r, zss = K.matvec(A, x, alpha=1.0, y=b.clone(), beta=-1.0)          # r = A x - b
    atr = K.rmatvec(A, r)
    rn = float(torch.sqrt(atr[n]))
    assert abs(rn / (0.1 * np.sqrt(m)) - 1) < 0.01                       # |r| ~ noise level
    assert float(torch.linalg.vector_norm(atr[:n])) <= 1e-9 * np.sqrt(m) * rn
    assert float(torch.linalg.vector_norm(x - x0)) <= 10 * 0.1 * np.sqrt(n / m) * np.sqrt(n)
    assert 20 <= log.iters <= 60 and log.errors[-1] <= 1e-9 * log.errors[0]
    assert log.passes_over_A <= log.iters + 4


@pytest.mark.parametrize("name", SPU_FIXTURES)
def test_spu1_matches_reference_fixture(rla, name):
    """Under-determined least squares min |y| s.t. A'y = c (SPU1, least_squares.py:425-494) with the
    reference's sketching operator replayed."""
    fx = load_golden(name)
    A, c = spu_problem_from_fixture(fx)
    m, n = A.shape
    S = sjlt_from_fixture(fx, int(float(fx["sf"]) * n), m)
    y, log = rla.SPU1(Replay(S), float(fx["sf"]))(dev(A), dev(c), float(fx["tol"]), int(fx["iter_lim"]), None)
    y = y.cpu().numpy()
    step = max(1, m // 64)
    assert np.linalg.norm(y[::step] - fx["y_probe"]) <= 1e-9 * float(fx["y_norm"])
    assert abs(np.linalg.norm(y) - float(fx["y_norm"])) <= 1e-9 * float(fx["y_norm"])
    assert abs(log.errors.size - fx["errors"].size) <= 1
    k = min(log.errors.size, fx["errors"].size)
    assert np.allclose(log.errors[:k], fx["errors"][:k], rtol=1e-5, atol=1e-10 * fx["errors"][0])
    # properties (test_underdet_least_squares.py:50-74): constraint satisfied, minimum norm
    y_opt = np.linalg.lstsq(A.T, c, rcond=None)[0]
    assert np.linalg.norm(A.T @ y - c) <= 1e-6 * np.linalg.norm(c)
    assert np.linalg.norm(y - y_opt) <= 1e-6 * np.linalg.norm(y_opt)


def test_pcss2_underdetermined_with_ridge(rla):
    """saddle.py:203-214 with delta > 0: x = (A'y - c)/delta solves the regularised saddle system."""
    rng = np.random.default_rng(12)
    m, n, delta = 1500, 60, 0.7
    A = rng.standard_normal((m, n)); c = rng.standard_normal(n)
    S = orc.sjlt_operator(4 * n, m, np.random.default_rng(2), 8)
    A_ske = np.vstack([S @ A, np.sqrt(delta) * np.eye(n)])
    M = orc.svd_right_precond(A_ske)[0]
    x_ref, y_ref, errs_ref = orc.pcss2_underdetermined(A, c, delta, 1e-12, 200, M, False)
    x, y, errs = rla.PcSS2()(dev(A), None, dev(c), delta, 1e-12, 200, dev(M), False, None)
    assert np.linalg.norm(y.cpu().numpy() - y_ref) <= 1e-9 * np.linalg.norm(y_ref)
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    assert abs(len(errs) - len(errs_ref)) <= 1


def test_spo_degenerate_inputs(rla):
    """b = 0 (LSQR's alfa*beta == 0 early exit, lsqr.py:392-395), a single column, iter_lim = 1, loose tol."""
    rng = np.random.default_rng(4)
    A = rng.standard_normal((400, 12))
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(dev(A), dev(np.zeros(400)), 0.0, 1e-12, 50, 1)
    assert float(x.abs().max()) == 0.0 and log.errors.shape == (2,)
    a1 = rng.standard_normal((300, 1)); b1 = rng.standard_normal(300)
    x, log = rla.SPO(rla.SkOpGA(), 4, 'qr')(dev(a1), dev(b1), 0.0, 1e-12, 50, 1)
    assert abs(float(x[0]) - float(a1[:, 0] @ b1 / (a1[:, 0] @ a1[:, 0]))) < 1e-12
    b = rng.standard_normal(400)
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'svd')(dev(A), dev(b), 0.0, 1e-12, 1, 1)       # one iteration only
    assert log.errors.size == 2
    x, log = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(dev(A), dev(b), 0.0, 0.5, 50, 1)         # loose tolerance stops early
    assert log.errors.size - 1 <= 3
    xs, _ = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(dev(A), dev(b), 0.0, 1e-12, 50, 1, logging=False)
    assert np.linalg.norm(xs.cpu().numpy() - np.linalg.lstsq(A, b, rcond=None)[0]) < 1e-10


def test_svd_right_precond_gram_route_and_fallback(rla):
    """preconditioning.py:70-79 for n >= 512: well-conditioned sketches take the Gram/eigh route (same
    M, U, sigma, Vh as the cuSOLVER SVD), ill-conditioned and rank-deficient ones fall back to it."""
    from parla_b200.comps import preconditioning as rpc
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 640
    X = torch.randn(3 * n, n, dtype=torch.float64, device="cuda", generator=g)
    for scale, expect_fast in ((None, True), (torch.logspace(0, -6, n, dtype=torch.float64, device="cuda"), False)):
        Xs = X if scale is None else X * scale
        assert (rpc._svd_via_gram(Xs) is not None) == expect_fast
        M, U, s, Vh = rpc.svd_right_precond(Xs)
        s_ref = torch.linalg.svdvals(Xs)
        assert float(((s.sort(descending=True)[0] - s_ref).abs() / s_ref).max()) <= 1e-10
        assert float(torch.linalg.norm((U * s) @ Vh - Xs) / torch.linalg.norm(Xs)) <= 1e-12
        eye = torch.eye(n, dtype=torch.float64, device="cuda")
        assert float(torch.linalg.norm(U.T @ U - eye)) <= 1e-9 and float(torch.linalg.norm(Vh @ Vh.T - eye)) <= 1e-9
        assert float(torch.linalg.norm(Xs @ M - U)) <= 1e-9 * n
    Xr = X.clone()
    Xr[:, -5:] = Xr[:, :5]                                   # rank n - 5: truncation must come from the true SVD
    assert rpc._svd_via_gram(Xr) is None
    M, U, s, Vh = rpc.svd_right_precond(Xr)
    assert M.shape == (n, n - 5) and s.numel() == n - 5
    # the driver: mode 'svd' (fast route) gives the same x as mode 'qr'
    A = torch.randn(20000, n, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(20000, dtype=torch.float64, device="cuda", generator=g)
    x_svd, log_svd = rla.SPO(rla.SkOpSJ(8), 4, 'svd')(A, b, 0.0, 1e-12, 100, 5)
    x_qr, log_qr = rla.SPO(rla.SkOpSJ(8), 4, 'qr')(A, b, 0.0, 1e-12, 100, 5)
    assert float(torch.linalg.vector_norm(x_svd - x_qr) / torch.linalg.vector_norm(x_qr)) <= 1e-10
    assert abs(log_svd.iters - log_qr.iters) <= 1
    assert np.allclose(log_svd.errors[:10], log_qr.errors[:10], rtol=1e-6)      # |M'A'r| is rotation invariant


def test_procedural_wrappers_run(rla):
    """sso1 / spo1 / spo3 / spu1 / sps (least_squares.py:108,193,200,419; saddlesys.py:77), svd1 / evd1 / evd2
    (svd.py:58, evd.py:16,99), qb / qb_b / qb_b_pe / rf1 / rs1 (qb.py:16,85,170; rangefinders.py:22; aware.py:9)
    and the small linear-algebra helpers of utils/linalg_wrappers.py."""
    from parla_b200.drivers import least_squares as ls, svd as dsvd, evd as devd
    from parla_b200.comps import qb as cqb, rangefinders as crf, preconditioning as cpc
    from parla_b200.comps.sketchers import aware
    from parla_b200.utils import linalg_wrappers as ulaw
    g = torch.Generator(device="cuda").manual_seed(9)
    rn = lambda *s: torch.randn(*s, dtype=torch.float64, device="cuda", generator=g)
    A, b, c = rn(2000, 30), rn(2000), rn(30)
    x_ls = torch.linalg.lstsq(A, b.reshape(-1, 1)).solution.reshape(-1)
    close = lambda x, y, t: float(torch.linalg.vector_norm(x - y) / torch.linalg.vector_norm(y)) <= t
    assert close(ls.spo1(A, b, 0.0, 1e-12, 100, 1)[0], x_ls, 1e-9)
    assert close(ls.spo3(A, b, 0.0, 1e-12, 100, 1, mode='chol')[0], x_ls, 1e-9)
    b_sig = A @ rn(30) + 0.01 * rn(2000)                          # small residual: sketch-and-solve is then accurate
    x_sig = torch.linalg.lstsq(A, b_sig.reshape(-1, 1)).solution.reshape(-1)
    assert close(ls.sso1(A, b_sig, 0.0, 1, sampling_factor=20)[0], x_sig, 1e-2)
    y, _ = ls.spu1(A, c, 1e-12, 100, 1)
    assert close(A.T @ y, c, 1e-9)
    x, yy, _ = rla.sps(A, b, c, 0.3, 1e-12, 100, 1, method='pcg')
    G = A.T @ A + 0.3 * torch.eye(30, dtype=torch.float64, device="cuda")
    assert close(x, torch.linalg.solve(G, A.T @ b - c), 1e-8)
    L = rn(400, 12) @ rn(12, 90)                                   # exact rank 12
    U, s, Vh = dsvd.svd1(L, 12, 4, 0.0, 2, 8, 3)
    assert close((U * s) @ Vh, L, 1e-10)
    for Q, B in (cqb.qb(2, L, 12, 3), cqb.qb_b(2, 4, False, L, 12, 0.0, 3), cqb.qb_b_pe(1, 4, L, 12, np.nan, 3)):
        assert close(Q @ B, L, 1e-9)
    assert crf.rf1(L, 12, 1, 3).shape == (400, 12) and aware.rs1(L, 12, 2, 3).shape == (90, 12)
    H = L.T @ L
    V, lam = devd.evd1(H, 12, 0.0, 4, 2, 8, 3)
    assert close((V * lam) @ V.T, H, 1e-9)
    V, lam = devd.evd2(H, 8, 3, 2, 3)
    assert V.shape == (90, 8) and bool((lam > 0).all())
    assert cpc.a_lift(A, 0.0) is A and cpc.a_lift(A, 2.0).shape == (2030, 30)
    M = rn(40, 25)
    Lf, Uf, P = ulaw.lupt(M)
    assert close(Lf @ Uf @ P.T, M, 1e-12)
    Lf, Uf, P = ulaw.lup(M)
    assert close(Lf @ Uf @ P, M, 1e-12)
    assert ulaw.lu_stabilize(M).shape == (40, 25)
    T = rn(40, 6)
    assert close(M @ ulaw.apply_pinv_on_left(T, M), M @ torch.linalg.pinv(M) @ T, 1e-9)
    Mw, Tr = rn(25, 40), rn(6, 40)                                 # wide operator: target @ pinv(operator) is 6 x 25
    assert close(ulaw.apply_pinv_on_right(Tr, Mw), Tr @ torch.linalg.pinv(Mw), 1e-9)
