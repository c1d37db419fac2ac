"""GPU (needs >= 2 devices, skipped otherwise): row-sharded SPO / SVD1 / EVD1 under torchrun + NCCL must
reproduce the single-GPU results of the same global problem (scripts/dist_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_row_sharded_drivers_match_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "scripts", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
