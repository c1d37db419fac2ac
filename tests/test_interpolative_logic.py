"""CPU: the interpolative-decomposition layer (comps/interpolative.py, drivers/interpolative.py) against the
oracle (which is pinned to the reference).  The column-pivoted QR and all index / coefficient logic are
device-agnostic tensor code, so they are checked here on CPU tensors; the only GPU kernel on this path, the
DMMA GEMM, is stubbed by torch.matmul FOR THIS TEST ONLY (the GPU run is tests/test_gpu_lowrank.py)."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

from oracle import parla_oracle as orc
import parla_b200 as rla
from parla_b200 import kernels as K
from parla_b200.comps.interpolative import qrcp, qrcp_osid


@pytest.fixture()
def cpu_gemm(monkeypatch):
    def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, out=None):
        C = alpha * ((A.T if transa else A) @ (B.T if transb else B))
        if out is None:
            return C
        out.copy_(C + beta * out)
        return out
    monkeypatch.setattr(K, "gemm", gemm)


def oracle_sketcher(num_pass):
    rs = orc.RS1(orc.SkOpGA(), num_pass, orc.orth, 1)
    return lambda A, k, rng: torch.from_numpy(np.ascontiguousarray(rs(A.numpy(), k, rng)))


def test_qrcp_matches_lapack_pivots():
    rng = np.random.default_rng(0)
    for trial in range(40):
        M, N = int(rng.integers(2, 40)), int(rng.integers(2, 120))
        rank = int(rng.integers(1, min(M, N) + 1))
        Y = rng.standard_normal((M, N)) if trial % 3 == 0 else rng.standard_normal((M, rank)) @ rng.standard_normal((rank, N))
        if trial % 5 == 0:
            Y = Y * np.logspace(0, -8, N)
        _, R, J = sla.qr(Y, mode='economic', pivoting=True)
        Rt, Jt = qrcp(torch.from_numpy(Y))
        kk = int(np.sum(np.abs(np.diag(R)) > 1e-10 * abs(R[0, 0])))          # numerically significant steps
        assert np.array_equal(J[:kk], Jt.numpy()[:kk])
        assert np.allclose(np.abs(np.diag(Rt.numpy()))[:kk], np.abs(np.diag(R))[:kk], rtol=1e-9)
        Q = torch.linalg.lstsq(Rt[:kk, :kk].T, torch.from_numpy(Y[:, Jt.numpy()[:kk]]).T).solution.T   # Y[:, J] R^-1
        assert np.allclose((Q.T @ Q).numpy(), np.eye(kk), atol=1e-6)
    X, Js = qrcp_osid(torch.from_numpy(Y), 1, 1)
    Xo, Jo = orc.qrcp_osid(Y, 1, 1)
    assert np.array_equal(Js.numpy(), Jo) and np.allclose(X.numpy(), Xo, atol=1e-12)
    with pytest.raises(ValueError):
        qrcp_osid(torch.from_numpy(Y), 1, 2)


@pytest.mark.parametrize("m,n,rank,k,over", [(100, 30, 30, 25, 4), (30, 100, 30, 27, 3), (100, 30, 5, 5, 1),
                                             (60, 60, 20, 12, 6)])
def test_id_and_cur_match_oracle(cpu_gemm, m, n, rank, k, over):
    A = orc.rand_low_rank(m, n, rank, np.random.default_rng(1))
    At = torch.from_numpy(A)
    sk_t, sk_o = oracle_sketcher(2), orc.RS1(orc.SkOpGA(), 2, orc.orth, 1)
    scale = lambda M: max(1.0, float(np.abs(M).max()))
    for axis in (0, 1):
        for T_, O_ in ((rla.OSID1, orc.OSID1), (rla.OSID2, orc.OSID2)):
            Mt, Pt = T_(sk_t)(At, k, over, axis, np.random.default_rng(3))
            Mo, Po = O_(sk_o)(A, k, over, axis, np.random.default_rng(3))
            assert np.array_equal(Pt.numpy(), Po)
            assert np.abs(Mt.numpy() - Mo).max() <= 1e-8 * scale(Mo)
            sub = Mt.numpy()[Po, :] if axis == 0 else Mt.numpy()[:, Po]       # test_osid.py:26-34
            assert np.linalg.norm(sub - np.eye(k)) < 1e-8
        assert np.array_equal(rla.ROCS1(sk_t)(At, k, over, axis, np.random.default_rng(3)).numpy(),
                              orc.ROCS1(sk_o)(A, k, over, axis, np.random.default_rng(3)))
    Zt, It, Xt, Jt = rla.TSID1(rla.OSID1(sk_t))(At, k, over, np.random.default_rng(3))
    Zo, Io, Xo, Jo = orc.TSID1(orc.OSID1(sk_o))(A, k, over, np.random.default_rng(3))
    assert np.array_equal(It.numpy(), Io) and np.array_equal(Jt.numpy(), Jo)
    assert np.abs(Zt.numpy() - Zo).max() <= 1e-8 * scale(Zo) and np.abs(Xt.numpy() - Xo).max() <= 1e-8 * scale(Xo)
    Jt, Ut, It = rla.CUR1(rla.OSID1(sk_t))(At, k, over, np.random.default_rng(3))
    Jo, Uo, Io = orc.CUR1(orc.OSID1(sk_o))(A, k, over, np.random.default_rng(3))
    assert np.array_equal(It.numpy(), Io) and np.array_equal(Jt.numpy(), Jo)
    cur_t = A[:, Jo] @ (Ut.numpy() @ A[Io, :])
    cur_o = A[:, Jo] @ (Uo @ A[Io, :])
    assert np.linalg.norm(cur_t - cur_o) <= 1e-8 * np.linalg.norm(A)
    if rank <= k:                                                             # test_cur.py: exact for rank <= k
        assert np.linalg.norm(A - cur_t) <= 1e-10 * np.linalg.norm(A)


def test_transposed_view_products(cpu_gemm):
    from parla_b200 import distla
    B = torch.randn(7, 5, dtype=torch.float64)
    S, Y = torch.randn(7, 3, dtype=torch.float64), torch.randn(5, 2, dtype=torch.float64)
    assert distla._is_transposed_view(B.T) and not distla._is_transposed_view(B)
    assert torch.allclose(distla.mm(B.T, S), B.T @ S) and torch.allclose(distla.mm_t(B.T, Y), B @ Y)
    assert torch.allclose(distla.mm(B, Y), B @ Y) and torch.allclose(distla.mm_t(B, S), B.T @ S)
