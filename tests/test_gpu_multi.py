"""Multi-GPU path on real GPUs (skipped on boxes with one GPU; the host-side logic of the sharding is covered on
the CPU with gloo in tests/test_host_logic.py)."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_step_with_peer_memory_reduction_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tests", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "reduction mode = peer" in res.stdout, res.stdout[-2000:]
