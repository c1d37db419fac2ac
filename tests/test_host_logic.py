"""CPU: host-side logic of the drivers (no kernels), and the world_size-2 gloo path of the
row-sharding helpers."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parla_b200 as rla
from parla_b200.comps.determiter.logging import SketchAndPrecondLog
from parla_b200.drivers.least_squares import dim_checks
from parla_b200.parallel import RowSharded, allreduce_, unwrap


def test_dim_checks_matches_reference_semantics():
    assert dim_checks(4, 1000, 50) == 200
    assert dim_checks(3.3, 1531, 77) == int(3.3 * 77)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert dim_checks(30, 100, 10) == 100          # clipped to n_rows, with a warning
        assert len(w) == 1
    with pytest.raises(AssertionError):
        dim_checks(4, 10, 50)                          # not tall


def test_log_wrap_up_layout():
    log = SketchAndPrecondLog()
    log.time_sketch, log.time_factor, log.time_presolve, log.time_iterate = 1.0, 2.0, 0.5, 4.0
    log.wrap_up(np.array([3.0, 2.0, 1.0]), 10.0)
    assert np.allclose(log.errors, [10.0, 3.0, 2.0, 1.0])
    assert np.allclose(log.times, [3.0, 3.5, 5.5, 7.5])
    log.wrap_up(np.float64(0.0), 1.0)                  # early-exit path hands a scalar
    assert log.errors.shape == (2,)


def test_public_names_and_exec_alias():
    for name in ("SPO", "SSO1", "SkOpGA", "SkOpSJ", "RS1", "RF1", "QB1", "QB2", "SVD1", "EVD1",
                 "gaussian_operator", "sjlt_operator", "RowSketcher", "RangeFinder", "QBDecomposer",
                 "OverLstsqSolver", "SVDecomposer", "EVDecomposer", "SketchAndPrecondLog"):
        assert hasattr(rla, name)
    for cls in (rla.SPO, rla.SSO1, rla.RS1, rla.RF1, rla.QB1, rla.QB2, rla.SVD1, rla.EVD1, rla.SkOpGA, rla.SkOpSJ):
        assert cls.exec is cls.__call__
    alg = rla.SPO(rla.SkOpSJ(8), 4, mode='qr')
    assert alg.mode == 'qr' and alg.sampling_factor == 4 and isinstance(alg.iterative_solver, rla.PcSS2)


def test_no_cpu_fallback():
    """CPU tensors are refused loudly, never routed to a host implementation."""
    from parla_b200 import kernels as K
    with pytest.raises(TypeError):
        K.gemm(torch.zeros(4, 4, dtype=torch.float64), torch.zeros(4, 4, dtype=torch.float64))
    with pytest.raises(TypeError):
        K.stream_pass(torch.zeros(8, 4, dtype=torch.float64), u=torch.zeros(8, dtype=torch.float64), flags=2)


def test_unknown_mode_raises_value_error_signature():
    import inspect
    sig = inspect.signature(rla.SPO.__call__)
    assert list(sig.parameters)[1:] == ["A", "b", "delta", "tol", "iter_lim", "rng", "logging"]
    sig = inspect.signature(rla.QB2.__init__)
    assert list(sig.parameters)[1:] == ["rf", "blk", "overwrite_a"]
    sig = inspect.signature(rla.RS1.__init__)
    assert list(sig.parameters)[1:] == ["sketch_op_gen", "num_pass", "stabilizer", "passes_per_stab"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m_local, n = 5, 3
        rng = np.random.default_rng(0)
        A = rng.standard_normal((world * m_local, n))
        u = rng.standard_normal(world * m_local)
        mine = slice(rank * m_local, (rank + 1) * m_local)
        shard = RowSharded.from_rank(torch.from_numpy(A[mine].copy()))
        assert shard.shape == (world * m_local, n) and shard.row_offset == rank * m_local
        local, off, group = unwrap(shard)
        # the per-iteration collective: partial [A^T u | |u|^2] summed over ranks
        part = torch.cat([local.T @ torch.from_numpy(u[mine]), torch.tensor([float(u[mine] @ u[mine])], dtype=torch.float64)])
        allreduce_(part, group)
        want = np.concatenate([A.T @ u, [u @ u]])
        ok = np.allclose(part.numpy(), want, rtol=1e-13)
        # a non-sharded tensor is passed through untouched
        t = torch.ones(2, dtype=torch.float64)
        allreduce_(t, unwrap(t)[2])
        ok = ok and torch.equal(t, torch.ones(2, dtype=torch.float64))
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


def test_row_sharded_allreduce_gloo_world2():
    world = 2
    out = mp.get_context("spawn").Array("i", [0] * world)
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port_, out))
             for port_ in [_free_port()] for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


class _LapackBlockQR:
    """CPU stand-in for the block-level QR kernels (torch.geqrf / ormqr): same contract, state kept in `ws`."""

    @staticmethod
    def workspace(device, d, n_layout):
        return {}

    @staticmethod
    def factor(panel, r0, c0, jb, tau_blk, k, n_layout, ws):
        a, tau = torch.geqrf(panel[r0:, c0:c0 + jb].clone())
        panel[r0:, c0:c0 + jb] = a
        tau_blk[:jb] = tau
        ws["a"], ws["tau"] = a, tau

    @staticmethod
    def apply(d, r0, jb, tau_blk, C, n_layout, ws):
        if C.shape[1]:
            C[r0:, :] = torch.ormqr(ws["a"], ws["tau"], C[r0:, :].clone(), left=True, transpose=True)


def _qr_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from parla_b200 import distla
        ok = True
        for d, n in ((700, 300), (600, 513), (1100, 1024)):       # partial last block, odd block counts per rank
            g = torch.Generator().manual_seed(5)
            W0 = torch.randn(d, n + 1, dtype=torch.float64, generator=g)
            W = W0.clone()
            tau = distla.geqrf_distributed(W, n, dist.group.WORLD, _backend=_LapackBlockQR)
            a_ref, tau_ref = torch.geqrf(W0[:, :n].clone())
            R_ref = torch.triu(a_ref[:n])
            qtb_ref = torch.ormqr(a_ref, tau_ref, W0[:, n:n + 1].clone(), left=True, transpose=True)[:n, 0]
            ok = ok and torch.allclose(torch.triu(W[:n, :n]), R_ref, rtol=0, atol=1e-11)
            ok = ok and torch.allclose(W[:n, n], qtb_ref, rtol=0, atol=1e-11)
            ok = ok and torch.allclose(tau, tau_ref, rtol=0, atol=1e-12)
        assert not distla.geqrf_distributed_ok(50000, 4096, dist.group.WORLD)      # too many rows for the block kernel
        assert not distla.geqrf_distributed_ok(8192, 300, dist.group.WORLD)        # too few blocks per rank
        assert distla.geqrf_distributed_ok(8192, 2048, dist.group.WORLD)
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


def test_column_distributed_qr_logic_gloo_world2():
    """Ownership, packing, panel broadcast and the final all-gather of distla.geqrf_distributed, with LAPACK standing
    in for the three block kernels (world_size 2, gloo, CPU)."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port_ = _free_port()
    procs = [ctx.Process(target=_qr_worker, args=(r, world, port_, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def test_saddle_and_srct_public_names():
    for name in ("SPS1", "SPS2", "sps", "SaddleSolver", "PcSS1", "PcSS2", "pcss1", "pcss2", "pcg", "SPU1",
                 "SkOpTC", "srct_operator", "generate_srct", "apply_srct", "QB3", "EVD2", "SkOpSS", "SkOpON", "SkOpIN",
                 "sparse_sign_operator", "orthonormal_operator", "sampling_operator"):
        assert hasattr(rla, name), name
    for cls in (rla.SPS1, rla.SPS2, rla.PcSS1, rla.PcSS2, rla.SkOpTC, rla.SPU1):
        assert cls.exec is cls.__call__
    a = rla.SPS1(rla.SkOpSJ(), 3)                          # saddlesys.py:120-129 defaults
    assert isinstance(a.iterative_solver, rla.PcSS1) and a.nystrom_strategy == 'left'
    assert isinstance(rla.SPS2(rla.SkOpSJ(), 3, None).iterative_solver, rla.PcSS2)
    with pytest.raises(ValueError):                        # saddlesys.py:84 (raised before any device work)
        rla.sps(None, None, None, 0.0, 1e-8, 10, 0, method='nope')
    with pytest.raises(NotImplementedError):               # abstract interfaces
        rla.SaddleSolver()(None, None, None, 0.0, 1e-8, 10, 0, True)
    with pytest.raises(NotImplementedError):
        rla.PrecondSaddleSolver()(None, None, None, 0.0, 1e-8, 10, None, False, None)


def test_srct_two_level_plan_choice():
    """The divisor m2 of m that minimises m2 + d / m2 (flops of the two DCT levels); None -> dense blocks."""
    from parla_b200.utils.sketching import SRCTOperator
    assert SRCTOperator.choose_m2(1 << 22, 8192) in (64, 128)
    assert SRCTOperator.choose_m2(100000, 10000) == 100
    assert SRCTOperator.choose_m2(1009, 132) is None       # prime row count: no factorisation
    assert SRCTOperator.choose_m2(2 * 1009, 100) is None   # only divisor in range would be 2 < 4
    m2 = SRCTOperator.choose_m2(6000, 300)
    assert 6000 % m2 == 0 and 4 <= m2 <= 512
    # level-1 table of the oracle-checked definition: cos/sin of pi * kappa * j2 / m2 in the row layout used
    m2 = 8
    kk = np.arange(m2 + 1)[:, None] * np.arange(m2)[None, :]
    assert np.allclose(np.cos(np.pi * (kk % (2 * m2)) / m2), np.cos(np.pi * kk / m2), atol=1e-15)


def test_procedural_wrappers_and_utils_exist():
    """Every module-level function of the reference's drivers/comps/utils has a counterpart with the same
    positional signature (parla/drivers/*.py, comps/*.py, utils/*.py)."""
    import inspect
    from parla_b200.drivers import least_squares as ls, svd as dsvd, evd as devd, saddlesys as dss
    from parla_b200.comps import qb as cqb, rangefinders as crf, preconditioning as cpc
    from parla_b200.comps.sketchers import aware
    from parla_b200.comps.determiter import saddle as dsad
    from parla_b200.utils import linalg_wrappers as ulaw
    want = {ls.sso1: ["A", "b", "delta", "rng", "sampling_factor", "vec_nnz", "lapack_driver"],
            ls.spo1: ["A", "b", "delta", "tol", "iter_lim", "rng", "sampling_factor", "vec_nnz"],
            ls.spo3: ["A", "b", "delta", "tol", "iter_lim", "rng", "sampling_factor", "vec_nnz", "mode"],
            ls.spu1: ["A", "c", "tol", "iter_lim", "rng", "sampling_factor", "vec_nnz"],
            dss.sps: ["A", "b", "c", "delta", "tol", "iter_lim", "rng", "sampling_factor", "vec_nnz", "method"],
            dsvd.svd1: ["A", "k", "over", "tol", "inner_num_pass", "block_size", "rng"],
            devd.evd1: ["A", "k", "tol", "over", "inner_num_pass", "block_size", "rng"],
            devd.evd2: ["A", "k", "over", "num_passes", "rng"],
            cqb.qb: ["num_passes", "A", "k", "rng"],
            cqb.qb_b: ["inner_num_pass", "blk", "overwrite_A", "A", "k", "tol", "rng"],
            cqb.qb_b_pe: ["num_passes", "blk", "A", "k", "tol", "rng"],
            crf.rf1: ["A", "k", "num_pass", "rng"],
            dsad.pcss1: ["A", "b", "c", "delta", "tol", "iter_lim", "R", "upper_tri", "z0"],
            dsad.pcss2: ["A", "b", "c", "delta", "tol", "iter_lim", "R", "upper_tri", "z0"],
            cpc.a_lift: ["A", "scale"], cpc.a_lift_precond: ["A", "delta", "R", "upper_tri", "k"]}
    for fn, names in want.items():
        assert list(inspect.signature(fn).parameters)[:len(names)] == names, fn.__name__
    assert list(inspect.signature(aware.rs1).parameters)[:4] == ["A", "k", "num_pass", "rng"]
    for name in ("orth", "lu_stabilize", "lupt", "lup", "apply_pinv_on_left", "apply_pinv_on_right"):
        assert callable(getattr(ulaw, name))


def test_fused_vector_phase_is_selected_only_where_it_applies():
    """comps/determiter/lsqr.py takes the fused path (pla_stream_pass_parts_f64 + pla_lsqr_fused_step_f64) only for a
    small dense row-major preconditioner on one GPU without ridge rows; everything else keeps the eager chain."""
    from parla_b200.comps.preconditioning import PrecondOperator
    from parla_b200.comps.determiter import lsqr as lsqr_mod
    A = torch.zeros(50, 12, dtype=torch.float64)
    M = torch.eye(12, dtype=torch.float64)
    cap = lsqr_mod.FUSED_MAX_ELEMS
    assert cap == 1 << 20 or "PLA_LSQR_FUSED_MAX_ELEMS" in os.environ
    assert PrecondOperator(A, 0.0, M, False).fusable(cap)
    assert not PrecondOperator(A, 0.0, M, True).fusable(cap)                     # triangular solves, not a dense M
    assert not PrecondOperator(A, 0.5, M, False).fusable(cap)                    # ridge rows
    assert PrecondOperator(A, 0.0, M[:, :7].contiguous(), False).fusable(cap)    # rank-truncated (svd mode)
    assert not PrecondOperator(A, 0.0, M.T.contiguous().T, False).fusable(cap)   # column-major M
    assert not PrecondOperator(A, 0.0, M, False).fusable(100)                    # too large for eight SMs
    wide = torch.zeros(4, 5000, dtype=torch.float64)
    assert not PrecondOperator(wide, 0.0, torch.zeros(5000, 200, dtype=torch.float64), False).fusable(cap)   # n_in > 4096
    assert PrecondOperator(wide[:, :3000], 0.0, torch.zeros(3000, 300, dtype=torch.float64), False).fusable(cap)
