"""Collects tests/multi_gpu_check.py under pytest when more than one CUDA device is visible:
every public API path under a 2-rank NCCL process group, bit-exact against the oracle on
every rank (skipped on a 1-GPU box)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(900)
def test_two_rank_api_parity_under_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           str(ROOT / "tests" / "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(ROOT), env=env, timeout=850)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi_gpu_check ok" in r.stdout
