"""parallel.ShardedRenderer on real GPUs over NCCL / symmetric memory (needs >= 2 GPUs in the box; skipped otherwise).
Runs tests/gpu_scripts/sharded_check.py under torchrun: every gather mode x frame dtype on a ragged clip, gathered clip
bit-equal to a local re-render of all ranks' chunks, host delivery, three frames against the oracle."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one host")
def test_sharded_renderer_on_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "gpu_scripts", "sharded_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    out = proc.stdout + proc.stderr
    assert "SHARDED_CHECK PASS" in out and proc.returncode == 0, out[-3000:]
