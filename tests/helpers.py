"""Shared test helpers (problem builders that mirror oracle/make_golden.py)."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def lsq_problem(m, n, seed, cond=1.0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((m, n))
    if cond != 1.0:
        A = A * np.logspace(0, np.log10(cond), n)
    x0 = rng.standard_normal(n)
    b = A @ x0 + 0.1 * rng.standard_normal(m)
    return A, b


def problem_from_fixture(fx):
    A, b = lsq_problem(int(fx["m"]), int(fx["n"]), int(fx["seed"]), float(fx["cond"]))
    assert digest(A) == str(fx["A_sha"]), "numpy's RNG stream differs from the one the fixture was generated with"
    if "b_sha" in fx:
        assert digest(b) == str(fx["b_sha"]), "b differs from the one the fixture was generated with"
    else:       # large problems: b = A @ x0 + noise is a host-BLAS gemv, equal across hosts only to round-off
        assert np.allclose(b[::4096], fx["b_probe"], rtol=0, atol=1e-12 * float(fx["b_norm"]) / np.sqrt(b.size))
        assert abs(np.linalg.norm(b) - float(fx["b_norm"])) <= 1e-13 * float(fx["b_norm"])
    return A, b


def sjlt_from_fixture(fx, d, m):
    """scipy CSC operator stored in (or regenerated for) a fixture."""
    import scipy.sparse as sps
    k = int(fx["vec_nnz"])
    if "S_rows" in fx:
        rows = fx["S_rows"].astype(np.int64)
        signs = fx["S_signs"].astype(np.float64)
    else:
        from oracle import parla_oracle as orc
        S = orc.sjlt_operator(d, m, np.random.default_rng(int(fx["rng_seed"])), k)
        r, s, _ = orc.sjlt_index_form(S)
        assert digest(r) + digest(s) == str(fx["S_sha"]), "regenerated SJLT differs from the fixture's"
        return S
    cols = np.repeat(np.arange(m), k)
    return sps.coo_matrix((signs.reshape(-1) / np.sqrt(k), (rows.reshape(-1), cols)), shape=(d, m)).tocsc()


SPO_FIXTURES = ["spo_sjlt_qr_600x40", "spo_sjlt_svd_600x40", "spo_sjlt_chol_600x40", "spo_sjlt_qr_ridge_600x40",
                "spo_sjlt_svd_ridge_600x40", "spo_gauss_qr_500x37", "spo_sjlt_qr_cond1e5_2000x64",
                "spo_sjlt_qr_odd_1531x77"]
SPO_RANKDEF_FIXTURES = ["spo_gauss_svd_rankdef_seed1", "spo_gauss_svd_rankdef_seed4"]
LOWRANK_BIG_FIXTURES = ["svd1_qb1_cfg4s_16384x2048", "svd1_qb2_cfg4s_16384x2048"]
SPU_FIXTURES = ["spu1_sjlt_800x50", "spu1_sjlt_2000x96"]
LOWRANK_FIXTURES = ["svd1_qb1_200x50", "svd1_qb2_200x50", "svd1_qb2_tol_200x50", "svd1_qb1_over_50x200", "evd1_qb1_120"]


SPS_FIXTURES = ["sps1_lin_1000x100", "sps1_lin_ridge_1000x100", "sps1_log_ridge_1000x100", "sps1_hiacc_500x50",
                "sps1_tiny_500x50", "sps1_nysleft_1000x100", "sps1_nysleft_ridge_1000x100", "sps1_nysright_1000x100",
                "sps1_nysright_ridge_1000x100", "sps2_lin_1000x100", "sps2_lin_ridge_1000x100",
                "sps2_log_ridge_1000x100", "sps2_hiacc_500x50"]


# PCG with a LOW-RANK (Nystrom) preconditioner and delta = 0 is not a contraction: the residual norm climbs
# back to ~0.5 |r0| mid-run and rounding differences grow ~10x per iteration.  The reference run against
# ITSELF with OPENBLAS_NUM_THREADS=1 vs 8 (measured in the build container, oracle == reference bit for bit)
# gives 77 vs 83 and 59 vs 51 iterations for these two fixtures, histories equal to 1e-6 only up to iteration
# 19 / 15, and x equal to 6e-12.  Parity for them is therefore: history prefix, iteration count within 20 %, x.
SPS_CHAOTIC = {"sps1_nysleft_1000x100": 12, "sps1_nysright_1000x100": 10}


def assert_history_close(h, h_ref, rtol=1e-6, chaotic_prefix=None, atol_rel=1e-9):
    """Iteration count and logged error history of an iterative solver against the reference's."""
    if chaotic_prefix is None:
        assert abs(h.size - h_ref.size) <= 1, (h.size, h_ref.size)
        k = min(h.size, h_ref.size)
        assert np.allclose(h[:k], h_ref[:k], rtol=rtol, atol=atol_rel * h_ref[0]), np.max(np.abs(h[:k] / h_ref[:k] - 1))
    else:
        assert abs(h.size - h_ref.size) <= 0.2 * h_ref.size, (h.size, h_ref.size)
        k = chaotic_prefix
        assert np.allclose(h[:k], h_ref[:k], rtol=rtol, atol=0), np.max(np.abs(h[:k] / h_ref[:k] - 1))


def saddle_problem_from_fixture(fx):
    """(A, b, c, x_opt) of the reference's saddle-system test problem (test_saddlesys.py:11-57)."""
    from oracle import parla_oracle as orc
    n, cond = int(fx["n"]), float(fx["cond"])
    if str(fx["kind"]) == "lin":
        spec = np.linspace(cond ** 0.5, cond ** -0.5, num=n)
    else:
        spec = np.logspace(np.log10(cond) / 2, -np.log10(cond) / 2, num=n)
    A, b, c, x_opt, _ = orc.saddle_problem(int(fx["m"]), n, spec, float(fx["delta"]),
                                           np.random.default_rng(int(fx["seed"])), float(fx["rhs_scale"]))
    assert digest(A) == str(fx["A_sha"]) and digest(b) == str(fx["b_sha"]) and digest(c) == str(fx["c_sha"]), \
        "numpy's RNG stream / LAPACK QR differs from the one the fixture was generated with"
    return A, b, c, x_opt


def sps_operator_from_fixture(fx):
    """The reference's sketching operator of a saddle fixture (SJLT stored; Gaussian regenerated + hash-checked)."""
    from oracle import parla_oracle as orc
    alg, m, n = str(fx["alg"]), int(fx["m"]), int(fx["n"])
    d = int(float(fx["sf"]) * n)
    if "S_rows" in fx:
        return sjlt_from_fixture(fx, d, m)
    shape = (n, d) if alg == "sps1_right" else (d, m)
    S = orc.gaussian_operator(shape[0], shape[1], np.random.default_rng(int(fx["rng_seed"])))
    assert digest(S) == str(fx["S_sha"]), "regenerated Gaussian operator differs from the fixture's"
    return S


def sps_algorithm(mod, fx, gen):
    """SPS1 / SPS2 of module ``mod`` (the oracle or parla_b200) configured as the fixture's case."""
    alg = str(fx["alg"])
    if alg == "sps2":
        return mod.SPS2(gen, float(fx["sf"]))
    a = mod.SPS1(gen, float(fx["sf"]))
    if alg in ("sps1_left", "sps1_right"):
        a.nystrom_strategy = alg.split("_")[1]
    return a


SRCT_FIXTURES = ["spo_srct_qr_600x40", "spo_srct_qr_ridge_600x40", "spo_srct_svd_600x40", "spo_srct_chol_600x40",
                 "spo_srct_qr_prime_1009x33", "spo_srct_qr_cond1e4_4096x96"]


def srct_from_fixture(fx, d, m):
    """The reference's SRCT operator of a fixture, as the oracle's operator object (carries sketch_data)."""
    from oracle import parla_oracle as orc
    e = fx["S_e_sign"].astype(np.float64) * float(fx["S_e_scale"])
    return orc.SrctOperator((d, m), (fx["S_r"].astype(np.int64), e, fx["S_perm"].astype(np.int64)))


QB3_FIXTURES = ["qb3_200x50", "qb3_tol_200x50", "qb3_wide_60x240"]
EVD2_FIXTURES = ["evd2_120_over0", "evd2_120_over5", "evd2_120_exact"]


def lowrank_matrix_from_fixture(fx):
    """Test matrix of a QB3 / EVD2 fixture (oracle generators, hash-checked)."""
    from oracle import parla_oracle as orc
    rng = np.random.default_rng(int(fx["seed"]))
    if str(fx["kind"]) == "evd2":
        n, rank = int(fx["n"]), int(fx["rank"])
        B0 = orc.rand_low_rank(n, rank, rank, rng)
        A = B0 @ B0.T
        A = 0.5 * (A + A.T)
    else:
        A = orc.exponent_spectrum(int(fx["m"]), int(fx["n"]), int(fx["rank"]), rng, 3.0)
    assert digest(A) == str(fx["A_sha"])
    return A


ID_FIXTURES = ["id_tall_100x30", "id_wide_30x100", "id_exact_100x30"]


def check_id_fixture(fx, A, lib, sk_op, to_np, wrap):
    """Run OSID1 / OSID2 / ROCS1 / TSID1 / CUR1 of ``lib`` (the oracle or parla_b200) on A and compare skeleton
    indices (exactly) and approximation errors (1e-8 relative to |A|) with a fixture of reference outputs."""
    k, over, fro = int(fx["k"]), int(fx["over"]), float(fx["A_fro"])
    Aw = wrap(A)
    for axis in (0, 1):
        for tag, cls in (("osid1", lib.OSID1), ("osid2", lib.OSID2)):
            M, P = cls(sk_op)(Aw, k, over, axis, np.random.default_rng(7))
            M, P = to_np(M), to_np(P)
            assert np.array_equal(P, fx[f"{tag}_ax{axis}_idx"]), (tag, axis)
            approx = M @ A[P, :] if axis == 0 else A[:, P] @ M
            assert abs(np.linalg.norm(A - approx) - float(fx[f"{tag}_ax{axis}_err"])) <= 1e-8 * fro
            sub = M[P, :] if axis == 0 else M[:, P]
            assert np.linalg.norm(sub - np.eye(k)) < 1e-8
        assert np.array_equal(to_np(lib.ROCS1(sk_op)(Aw, k, over, axis, np.random.default_rng(7))),
                              fx[f"rocs1_ax{axis}_idx"])
    Z, Is, X, Js = (to_np(v) for v in lib.TSID1(lib.OSID1(sk_op))(Aw, k, over, np.random.default_rng(7)))
    assert np.array_equal(Is, fx["tsid_Is"]) and np.array_equal(Js, fx["tsid_Js"])
    assert abs(np.linalg.norm(A - Z @ A[Is, :][:, Js] @ X) - float(fx["tsid_err"])) <= 1e-8 * fro
    Js, U, Is = (to_np(v) for v in lib.CUR1(lib.OSID1(sk_op))(Aw, k, over, np.random.default_rng(7)))
    assert np.array_equal(Is, fx["cur_Is"]) and np.array_equal(Js, fx["cur_Js"])
    assert abs(np.linalg.norm(A - A[:, Js] @ (U @ A[Is, :])) - float(fx["cur_err"])) <= 1e-8 * fro


def id_matrix_from_fixture(fx):
    from oracle import parla_oracle as orc
    A = orc.rand_low_rank(int(fx["m"]), int(fx["n"]), int(fx["rank"]), np.random.default_rng(int(fx["seed"])))
    assert digest(A) == str(fx["A_sha"])
    return A


class Replay:
    """sketch_op_gen that hands back a prerecorded operator (reference S replayed on the GPU path)."""

    def __init__(self, S):
        self.S = S

    def __call__(self, n_rows, n_cols, rng):
        assert self.S.shape == (n_rows, n_cols)
        return self.S


def spu_problem_from_fixture(fx):
    rng = np.random.default_rng(int(fx["seed"]))
    m, n = int(fx["m"]), int(fx["n"])
    A = rng.standard_normal((m, n)) * np.logspace(0, np.log10(float(fx["cond"])), n)
    c = rng.standard_normal(n)
    assert digest(A) == str(fx["A_sha"]) and digest(c) == str(fx["c_sha"])
    return A, c
