"""D-sharded MoDE path on >= 2 GPUs: launches tests/check_sharded.py under torchrun, one rank
per GPU over NCCL.  Skipped on boxes with a single GPU (the gloo world_size-2 tests in test_parallel_cpu.py cover the
host logic everywhere)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_stage_and_net_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "check_sharded.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0 and "SHARDED_CHECK_OK" in r.stdout, tail
