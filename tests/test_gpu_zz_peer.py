"""GPU: the cross-rank sum fused into the streaming pass's reduce kernel (pla_stream_pass_peer_f64, NVLink peer memory /
CUDA IPC).  Runs last (file name): the two-process case shares cuda:0 between two ranks, which a box in exclusive-process
compute mode does not allow -- that case is then skipped, not failed."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def K():
    from parla_b200 import kernels
    return kernels


# ------------------------------------------------------------------ fused reduce + cross-rank sum (NVLink peer memory)
def test_stream_pass_peer_exchange_with_itself(K):
    """world = 1: the reduce kernel pushes its totals into its own exchange buffer and reads them back; the result
    must be the plain pass bit for bit, call after call (slot reuse across epochs, different widths)."""
    from parla_b200.parallel import PeerComm
    comm = PeerComm(None, torch.device("cuda", torch.cuda.current_device()), 8193)
    assert comm.ok and comm.world == 1
    try:
        for trial, (m, n) in enumerate([(3000, 64), (5000, 500), (2048, 2048), (1000, 8192), (3000, 64), (777, 333)]):
            rng = np.random.default_rng(trial)
            A, w, u0 = dev(rng.standard_normal((m, n))), dev(rng.standard_normal(n)), dev(rng.standard_normal(m))
            for flags in (K.PASS_DOT | K.PASS_AXPY, K.PASS_AXPY, K.PASS_DOT):
                kw = dict(w=w if flags & K.PASS_DOT else None, sa=0.5, su=-1.0, flags=flags)
                ul, uf = u0.clone(), u0.clone()
                z_loc = K.stream_pass(A, u=ul, **kw)
                z_fused = K.stream_pass(A, u=uf, comm=comm, **kw)
                assert torch.equal(z_loc, z_fused) and torch.equal(ul, uf)
        assert comm.epoch == 18
    finally:
        torch.cuda.synchronize()
        comm.close()


def test_stream_pass_peer_exchange_two_processes():
    """Two ranks (two processes) exchange through CUDA IPC mappings; both sit on cuda:0, so this runs on a one-GPU
    box.  scripts/peer_check.py compares the fused result with the rank-ordered sum of the local results, bit for bit."""
    import socket, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "scripts", "peer_check.py"), "--same-gpu"]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=420)
    if out.returncode != 0 and any(t in out.stderr for t in ("exclusive", "busy or unavailable", "all CUDA-capable devices are busy")):
        pytest.skip("the GPU does not admit a second process (exclusive compute mode)")
    assert out.returncode == 0 and "PEER_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
