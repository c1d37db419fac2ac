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
    assert digest(A) == str(fx["A_sha"]) and digest(b) == str(fx["b_sha"]), \
        "numpy's RNG stream differs from the one the fixture was generated with"
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
SPU_FIXTURES = ["spu1_sjlt_800x50", "spu1_sjlt_2000x96"]
LOWRANK_FIXTURES = ["svd1_qb1_200x50", "svd1_qb2_200x50", "svd1_qb2_tol_200x50", "svd1_qb1_over_50x200", "evd1_qb1_120"]


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
