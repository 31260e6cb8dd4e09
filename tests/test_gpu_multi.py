"""Multi-GPU exchange on real devices (needs >= 2 GPUs; skipped on the single-GPU test box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_fused_exchange_matches_collective_route_and_exact_merge():
    probe = os.path.join(os.path.dirname(__file__), "_sharded_probe.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", probe]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "ok=True" in r.stdout and "fused" in r.stdout
