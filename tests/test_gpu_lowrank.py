"""GPU: RS1 / RF1 / QB1 / QB2 / SVD1 / EVD1 against the golden fixtures (reference outputs, numpy
Gaussian test matrices replayed) and the reference's own property tests
(parla/tests/test_comps/test_qb.py:91-121, test_drivers/test_lowrank/test_svd.py:28-62, test_evd.py:35-63)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import parla_oracle as orc
from tests.helpers import (EVD2_FIXTURES, ID_FIXTURES, LOWRANK_BIG_FIXTURES, LOWRANK_FIXTURES, QB3_FIXTURES,
                           check_id_fixture, digest, id_matrix_from_fixture, load_golden, lowrank_matrix_from_fixture)

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


@pytest.fixture(scope="module")
def rla():
    import parla_b200
    return parla_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_matrix(fx):
    m, n, rank = (int(fx[q]) for q in ("m", "n", "rank"))
    rng = np.random.default_rng(int(fx["seed"]))
    if bool(fx["evd"]):
        B0 = orc.rand_low_rank(m, rank, rank, rng)
        A = B0 @ B0.T
        A = 0.5 * (A + A.T)
    else:
        A = orc.exponent_spectrum(m, n, rank, rng, 2.0)
    assert digest(A) == str(fx["A_sha"])
    return A


@pytest.mark.parametrize("name", LOWRANK_FIXTURES)
def test_lowrank_matches_reference_fixture(rla, name):
    fx = load_golden(name)
    A = build_matrix(fx)
    k, tol, over = int(fx["k"]), float(fx["tol"]), int(fx["over"])
    # numpy Gaussian test matrices == the reference's (oracle generator is bit-identical), replayed on GPU
    rs = rla.RS1(orc.SkOpGA(), int(fx["num_pass"]), rla.orth, 1)
    rf = rla.RF1(rs)
    qb = rla.QB1(rf) if int(fx["blk"]) < 0 else rla.QB2(rf, int(fx["blk"]), False)
    Ad = dev(A)
    if bool(fx["evd"]):
        V, lam = rla.EVD1(qb)(Ad, k, tol, over, np.random.default_rng(7))
        V, spec = V.cpu().numpy(), lam.cpu().numpy()
        approx = (V * spec) @ V.T
        assert np.linalg.norm(V.T @ V - np.eye(V.shape[1])) < 1e-10
    else:
        U, s, Vh = rla.SVD1(qb)(Ad, k, tol, over, np.random.default_rng(7))
        U, spec, Vh = U.cpu().numpy(), s.cpu().numpy(), Vh.cpu().numpy()
        approx = (U * spec) @ Vh
        assert np.all(spec >= 0) and np.all(np.diff(spec) <= 0)
        assert np.linalg.norm(U.T @ U - np.eye(U.shape[1])) < 1e-8 and np.linalg.norm(Vh @ Vh.T - np.eye(Vh.shape[0])) < 1e-8
    assert torch.equal(Ad.cpu(), torch.from_numpy(A))                      # test_unchanged_A
    assert spec.shape == fx["spec"].shape
    assert np.max(np.abs(spec - fx["spec"])) <= 1e-10 * np.max(np.abs(fx["spec"]))
    sr, sc = max(1, A.shape[0] // 16), max(1, A.shape[1] // 16)
    assert np.max(np.abs(approx[::sr, ::sc] - fx["approx_probe"])) <= 1e-10 * float(fx["approx_fro"])
    assert abs(np.linalg.norm(A - approx) - float(fx["err_fro"])) <= 1e-10 * float(fx["approx_fro"])


@pytest.mark.parametrize("name", LOWRANK_BIG_FIXTURES)
def test_lowrank_cfg4_scaled_parity(rla, name):
    """BASELINE.json configs[3] scaled as SURVEY.md 8(d) prescribes: 2^14 x 2^11, decaying spectrum, k = 128, two
    power iterations (RS1), QB1 and QB2(blk = 32); SVD1 against (1) the REFERENCE's spectrum / approximation
    (fixture) and (2) the oracle run live on the same matrix with the same numpy Gaussian test matrices.
    Tolerances: s to 1e-10 s_0, |U s V' - (U s V')_ref|_F <= 1e-10 |A|_F."""
    fx = load_golden(name)
    m, n, k = int(fx["m"]), int(fx["n"]), int(fx["k"])
    A = orc.exponent_spectrum(m, n, int(fx["rank"]), np.random.default_rng(int(fx["seed"])), float(fx["spectrum_param"]))
    fro = float(fx["A_fro"])
    assert abs(np.linalg.norm(A) - fro) <= 1e-13 * fro
    assert np.max(np.abs(A[::m // 16, ::n // 16] - fx["A_probe"])) <= 1e-13 * np.max(np.abs(fx["A_probe"]))
    blk, tol = int(fx["blk"]), float(fx["tol"])

    def build(lib, orth):
        rf = lib.RF1(lib.RS1(orc.SkOpGA(), int(fx["num_pass"]), orth, 1))
        return lib.SVD1(lib.QB1(rf) if blk < 0 else lib.QB2(rf, blk, False))

    U, s, Vh = build(rla, rla.orth)(dev(A), k, tol, 0, np.random.default_rng(7))
    U, s, Vh = U.cpu().numpy(), s.cpu().numpy(), Vh.cpu().numpy()
    assert s.shape == fx["spec"].shape and np.max(np.abs(s - fx["spec"])) <= 1e-10 * fx["spec"][0]
    approx = (U * s) @ Vh
    assert np.max(np.abs(approx[::m // 16, ::n // 16] - fx["approx_probe"])) <= 1e-10 * float(fx["approx_fro"])
    assert abs(np.linalg.norm(A - approx) - float(fx["err_fro"])) <= 1e-10 * fro
    assert np.linalg.norm(U.T @ U - np.eye(k)) < 1e-9 and np.linalg.norm(Vh @ Vh.T - np.eye(k)) < 1e-9
    Uo, so, Vho = build(orc, orc.orth)(A, k, tol, 0, np.random.default_rng(7))            # live oracle, full matrices
    assert np.max(np.abs(s - so)) <= 1e-10 * so[0]
    assert np.linalg.norm(approx - (Uo * so) @ Vho) <= 1e-10 * fro


@pytest.mark.parametrize("shape,rank", [((200, 50), 15), ((50, 200), 15), ((200, 50), 50)])
@pytest.mark.parametrize("which", ["qb1", "qb2", "qb2_sjlt"])
def test_qb_properties_native_operators(rla, shape, rank, which):
    """test_qb.py:91-121 with the Philox-native Gaussian / SJLT test matrices."""
    for seed in (1, 4, 15):
        A = orc.rand_low_rank(shape[0], shape[1], rank, np.random.default_rng(seed))
        gen = rla.SkOpSJ(vec_nnz=4) if which.endswith("sjlt") else rla.SkOpGA()
        rf = rla.RF1(rla.RS1(gen, 1, rla.orth, 1))
        qb = rla.QB1(rf) if which == "qb1" else rla.QB2(rf, 4, False)
        Ad = dev(A)
        Q, B = qb(Ad, rank, 0.0 if which != "qb1" else np.nan, np.random.default_rng(seed))
        Q, B = Q.cpu().numpy(), B.cpu().numpy()
        assert Q.shape == (shape[0], rank) and B.shape == (rank, shape[1])
        assert np.linalg.norm(Q.T @ Q - np.eye(rank)) <= 1e-8             # test_valid_onb
        assert np.linalg.norm(A - Q @ B) <= 1e-8                          # test_exact
        assert np.linalg.norm(B - Q.T @ A) <= 1e-8                        # test_exact_B
        assert torch.equal(Ad.cpu(), torch.from_numpy(A))                 # test_unchanged_A


def test_orth_tall_cholqr2_and_fallback(rla):
    """orth of a tall operand: CholeskyQR2 when well conditioned (orthonormal to machine precision, same range),
    Householder when the first round's orthogonality defect says otherwise (ill-conditioned / rank-deficient)."""
    from parla_b200 import distla
    g = torch.Generator(device="cuda").manual_seed(4)
    m, k = 1 << 16, 96
    Y = torch.randn(m, k, dtype=torch.float64, device="cuda", generator=g) * torch.logspace(0, -3, k, dtype=torch.float64, device="cuda")
    assert distla._cholqr2(Y, None) is not None
    Q = rla.orth(Y)
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    assert float(torch.linalg.norm(Q.T @ Q - eye)) < 1e-13
    assert float(torch.linalg.norm(Y - Q @ (Q.T @ Y)) / torch.linalg.norm(Y)) < 1e-13          # same range
    # (a bad COLUMN SCALING alone does not hurt Cholesky QR; an ill-conditioned mixing of the columns does)
    Qk = torch.linalg.qr(torch.randn(k, k, dtype=torch.float64, device="cuda", generator=g)).Q
    Y0 = torch.randn(m, k, dtype=torch.float64, device="cuda", generator=g)
    Yb = Y0 @ ((Qk * torch.logspace(0, -11, k, dtype=torch.float64, device="cuda")) @ Qk.T)       # cond ~ 1e11
    assert distla._cholqr2(Yb, None) is None
    Qb = rla.orth(Yb)
    assert float(torch.linalg.norm(Qb.T @ Qb - eye)) < 1e-12
    Yr = Y.clone()
    Yr[:, -3:] = Yr[:, :3]                                                                        # rank k - 3
    assert distla._cholqr2(Yr, None) is None
    Qr = rla.orth(Yr)
    assert float(torch.linalg.norm(Qr.T @ Qr - eye)) < 1e-12
    # the driver on a tall matrix (fast path inside RS1 / RF1 / QB1) against the live oracle
    A = orc.exponent_spectrum(1 << 16, 256, 200, np.random.default_rng(8), 20.0)
    alg = lambda lib, orth: lib.SVD1(lib.QB1(lib.RF1(lib.RS1(orc.SkOpGA(), 2, orth, 1))))
    U, s, Vh = alg(rla, rla.orth)(dev(A), 48, np.nan, 0, np.random.default_rng(7))
    Uo, so, Vho = alg(orc, orc.orth)(A, 48, np.nan, 0, np.random.default_rng(7))
    assert np.max(np.abs(s.cpu().numpy() - so)) <= 1e-10 * so[0]
    approx = (U * s) @ Vh
    assert np.linalg.norm(approx.cpu().numpy() - (Uo * so) @ Vho) <= 1e-10 * np.linalg.norm(A)


def test_qb2_tolerance_and_overwrite(rla):
    A = orc.exponent_spectrum(300, 120, 100, np.random.default_rng(3), 4.0)
    rf = rla.RF1(rla.RS1(rla.SkOpGA(), 0, rla.orth, 1))
    Ad = dev(A)
    Q, B = rla.QB2(rf, 8, True)(Ad, 120, 1e-3, 5)                            # overwrite_a: A is deflated in place
    Q, B = Q.cpu().numpy(), B.cpu().numpy()
    assert np.linalg.norm(A - Q @ B) <= 1e-3 * np.linalg.norm(A) and Q.shape[1] < 120
    assert np.linalg.norm(Ad.cpu().numpy() - (A - Q @ B)) <= 1e-10 * np.linalg.norm(A)
    with warnings.catch_warnings(record=True) as w:                          # k clipped with a warning
        warnings.simplefilter("always")
        Q, B = rla.QB2(rf, 50, False)(dev(A[:, :30]), 99, 0.0, 5)
        assert Q.shape[1] == 30 and any("target rank" in str(i.message) for i in w)
    with pytest.raises(AssertionError):
        rla.QB1(rf)(dev(A), 0, np.nan, 1)


def test_rs1_power_iteration_improves_alignment(rla):
    """test_aware.py:45-106 in spirit: more passes => S aligns with the dominant right singular space."""
    rng = np.random.default_rng(0)
    spec = np.concatenate([np.linspace(1.0, 0.8, 5), np.logspace(-1, -3, 55)])          # gap after the 5th value
    A, U, s, Vt = orc.rand_low_rank(400, 60, spec, rng, factors=True)
    k = 5
    errs = []
    for num_pass in (0, 1, 2, 3, 4, 6):
        S = rla.RS1(rla.SkOpGA(), num_pass, rla.orth, 1)(dev(A), k, 11).cpu().numpy()
        assert S.shape == (60, k)
        Qs = np.linalg.qr(S)[0]
        errs.append(np.linalg.norm(Vt[:k] - (Vt[:k] @ Qs) @ Qs.T))
    assert errs[-1] < 1e-2 * errs[0] and all(b <= a * 1.5 for a, b in zip(errs, errs[1:]))


def test_svd1_fixed_precision_and_evd_indefinite(rla):
    A = orc.exponent_spectrum(256, 128, 100, np.random.default_rng(2), 3.0)
    qb = rla.QB2(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1)), 16, False)
    U, s, Vh = rla.SVD1(qb)(dev(A), 128, 1e-6, 0, 1)
    approx = (U * s) @ Vh
    assert float(torch.linalg.norm(dev(A) - approx)) <= 1e-6 * np.linalg.norm(A)
    # symmetric indefinite EVD: eigenvalues ordered by decreasing magnitude
    rng = np.random.default_rng(4)
    Q0 = np.linalg.qr(rng.standard_normal((150, 150)))[0]
    lam0 = np.concatenate([[9, -8, 7, -6, 5], 1e-3 * rng.standard_normal(145)])
    H = (Q0 * lam0) @ Q0.T
    H = 0.5 * (H + H.T)
    V, lam = rla.EVD1(rla.QB1(rla.RF1(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1))))(dev(H), 5, np.nan, 5, 3)
    lam = lam.cpu().numpy()
    assert np.allclose(lam, [9, -8, 7, -6, 5], atol=1e-4) and V.shape == (150, 5)


@pytest.mark.parametrize("name", QB3_FIXTURES + EVD2_FIXTURES)
def test_qb3_evd2_match_reference_fixture(rla, name):
    """QB3 (comps/qb.py:484-598) and EVD2 (drivers/evd.py:290-381), numpy Gaussian test matrices replayed."""
    fx = load_golden(name)
    A = lowrank_matrix_from_fixture(fx)
    Ad = dev(A)
    if str(fx["kind"]) == "qb3":
        alg = rla.QB3(rla.RS1(orc.SkOpGA(), 0, rla.orth, 1), int(fx["blk"]))
        Q, B = alg(Ad, int(fx["k"]), float(fx["tol"]), np.random.default_rng(7))
        Q, B = Q.cpu().numpy(), B.cpu().numpy()
        assert Q.shape[1] == int(fx["qb_cols"]) and B.shape == (Q.shape[1], A.shape[1])
        assert np.linalg.norm(Q.T @ Q - np.eye(Q.shape[1])) < 1e-10              # test_valid_onb
        assert np.linalg.norm(B - Q.T @ A) <= 1e-8 * np.linalg.norm(A)            # test_exact_B
        approx = Q @ B
    else:
        alg = rla.EVD2(rla.RS1(orc.SkOpGA(), 1, rla.orth, 1))
        V, lam = alg(Ad, int(fx["k"]), np.nan, int(fx["over"]), np.random.default_rng(7))
        V, lam = V.cpu().numpy(), lam.cpu().numpy()
        assert lam.shape == fx["spec"].shape and np.all(lam > 0)
        assert np.max(np.abs(lam - fx["spec"])) <= 1e-9 * np.max(fx["spec"])
        assert np.linalg.norm(V.T @ V - np.eye(V.shape[1])) < 1e-10
        approx = (V * lam) @ V.T
    assert torch.equal(Ad.cpu(), torch.from_numpy(A))                             # test_unchanged_A
    sr, sc = max(1, A.shape[0] // 16), max(1, A.shape[1] // 16)
    assert np.max(np.abs(approx[::sr, ::sc] - fx["approx_probe"])) <= 1e-9 * float(fx["approx_fro"])
    assert abs(np.linalg.norm(A - approx) - float(fx["err_fro"])) <= 1e-9 * float(fx["approx_fro"])


def test_qb3_evd2_interface_errors(rla):
    A = torch.randn(40, 30, dtype=torch.float64, device="cuda")
    with pytest.raises(AssertionError):
        rla.QB3(rla.RS1(rla.SkOpGA(), 0, rla.orth, 1), 4)(A, 30, np.nan, 0)       # needs k < min(A.shape)
    with pytest.raises(RuntimeError):
        rla.QB3(lambda A_, k_, rng_: "not a matrix", 4)(A, 5, np.nan, 0)          # qb.py:568-574
    H = A.T @ A
    with pytest.raises(AssertionError):
        rla.EVD2(rla.RS1(rla.SkOpGA(), 1, rla.orth, 1))(H, 30, np.nan, 0, 0)
    with pytest.warns(UserWarning):
        rla.EVD2(rla.RS1(rla.SkOpGA(), 1, rla.orth, 1))(H, 5, 1e-3, 2, 0)
    Q, B = rla.QB3(rla.RS1(rla.SkOpGA(), 2, rla.orth, 1), 8)(A, 20, np.nan, 0)    # native operators
    assert float(torch.linalg.norm(Q.T @ Q - torch.eye(20, dtype=torch.float64, device="cuda"))) < 1e-10


@pytest.mark.parametrize("name", ID_FIXTURES)
def test_interpolative_matches_reference_fixture(rla, name):
    """OSID1 / OSID2 / ROCS1 / TSID1 / CUR1 on the device (numpy Gaussian test matrices replayed): skeleton indices
    equal to the reference's, approximation errors equal to 1e-8 |A|."""
    import types
    fx = load_golden(name)
    A = id_matrix_from_fixture(fx)
    sk_op = rla.RS1(orc.SkOpGA(), int(fx["num_pass"]), rla.orth, 1)
    lib = types.SimpleNamespace(OSID1=rla.OSID1, OSID2=rla.OSID2, ROCS1=rla.ROCS1, TSID1=rla.TSID1, CUR1=rla.CUR1)
    check_id_fixture(fx, A, lib, sk_op, lambda t: t.cpu().numpy(), dev)


def test_interpolative_procedural_and_errors(rla):
    from parla_b200.drivers import interpolative as did
    from parla_b200.comps import interpolative as cid
    A = torch.randn(300, 12, dtype=torch.float64, device="cuda") @ torch.randn(12, 80, dtype=torch.float64, device="cuda")
    Z, Is = did.osid1(A, 12, 3, 2, 0, 1)
    assert float(torch.linalg.norm(Z @ A[Is, :] - A) / torch.linalg.norm(A)) < 1e-10
    X, Js = did.osid2(A, 12, 3, 1, 1, 1)
    assert float(torch.linalg.norm(A[:, Js] @ X - A) / torch.linalg.norm(A)) < 1e-10
    Z, Is, X, Js = did.tsid1(A, 12, 3, 2, 1)
    assert float(torch.linalg.norm(Z @ A[Is, :][:, Js] @ X - A) / torch.linalg.norm(A)) < 1e-9
    Js, U, Is = did.cur1(A, 12, 3, 2, 1)
    assert float(torch.linalg.norm(A[:, Js] @ (U @ A[Is, :]) - A) / torch.linalg.norm(A)) < 1e-9
    assert cid.rocs1(A, 12, 3, 2, 0, 1).numel() == 12
    with pytest.raises(ValueError):
        did.OSID1(rla.RS1(rla.SkOpGA(), 0, rla.orth, 1))(A, 5, 2, 2, 0)
