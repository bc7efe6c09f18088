"""D-sharded MoDE path.

  * on ANY box with one GPU: two ranks share cuda:0 (two processes under torchrun) and run the D-sharded headline block
    -- halo exchange of the operand and of dy, BatchNorm statistics over owned planes, gradient sum -- against the
    UNSHARDED oracle (oracle/mode_torch.py on the whole volume): once with the exchange steps as stores into the other
    rank's memory over CUDA IPC (peer.PeerComm, the product path, kernel-fused push / gather included) and once through
    torch.distributed (gloo with host staging: NCCL refuses two ranks on one device).
  * on boxes with >= 2 GPUs additionally: the same over NCCL and over NVLink peer memory with one GPU per rank, and the
    two-conv stage / whole-U-Net sharding check (tests/check_sharded.py).
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _torchrun(script, args, nproc, port, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)] + args
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-30:])
    return r, tail


@pytest.mark.parametrize("comm", ["peer", "gloo"])
def test_sharded_block_two_ranks_on_one_gpu_vs_oracle(comm):
    r, tail = _torchrun("check_sharded_block.py", ["--comm", comm, "--same-device"], 2, 29561 if comm == "peer" else 29562)
    assert r.returncode == 0 and "SHARDED_BLOCK_OK" in r.stdout, tail


@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_sharded_block_one_gpu_per_rank_vs_oracle(comm):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    r, tail = _torchrun("check_sharded_block.py", ["--comm", comm], min(n, 4), 29563 if comm == "peer" else 29564)
    assert r.returncode == 0 and "SHARDED_BLOCK_OK" in r.stdout, tail


def test_sharded_stage_and_net_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r, tail = _torchrun("check_sharded.py", [], 2, 29533, timeout=600)
    assert r.returncode == 0 and "SHARDED_CHECK_OK" in r.stdout, tail


def test_module_on_second_device_while_first_is_current():
    """The reference only moves tensors (`.to(cuda:gpu_ids[0])`, fnet_model.py:53) and never calls set_device: a module on
    cuda:1 must run there -- kernels, stream, tensor maps, error flag -- while the process's current device stays 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from repmode_b200.nn_modules import MoDEConv
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    m0 = MoDEConv(5, 12, 32, 32).train()
    m1 = MoDEConv(5, 12, 32, 32).train()
    m1.load_state_dict(m0.state_dict())
    m0, m1 = m0.to("cuda:0"), m1.to("cuda:1")
    x = torch.randn(1, 32, 4, 16, 16)
    t = torch.tensor([3])
    outs = []
    for m, dev in ((m0, "cuda:0"), (m1, "cuda:1")):
        xi = x.to(dev).requires_grad_(True)
        y = m(xi, t.to(dev))
        y.square().sum().backward()
        outs.append((y.detach().cpu(), xi.grad.cpu(), m.gate.weight.grad.cpu()))
        assert torch.cuda.current_device() == 0
    for a, b in zip(*outs):
        assert torch.equal(a, b)
