"""CPU, gloo, world_size 2 and 4: the HOST logic of the D-sharded whole-U-Net forward (repmode_b200/sharded.py:
sharded_net_forward -- which levels run as slabs, the 4-plane halo exchange per two-conv stage, the gather / re-slice
transitions around the levels thinner than the halo, global BatchNorm statistics over owned planes) against the UNSHARDED
oracle (oracle/mode_torch.py on the whole volume).

The CUDA entry points the host logic calls (functional.mode_conv / down_conv_bn_relu / up_conv_bn_relu) are replaced by
CPU restatements built from the oracle's own pieces that honour the ShardSpec contract of the kernels (statistics over
the OWNED planes, all-reduced over the ranks; planes outside the global volume forced to zero): what is under test is the
slab bookkeeping, which is the same code on the GPU.  world = 4 has interior ranks (a neighbour on both sides), which the
2-rank tests never see."""
import argparse
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import mode_torch as otc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _bn_relu_spec(y, bn, shard):
    """Train-mode BatchNorm + ReLU with the kernels' ShardSpec semantics (bn.cu / conv epilogue statistics)."""
    w, b = bn[0], bn[1]
    if shard is None:
        return F.relu(F.batch_norm(y, None, None, w, b, True, 0.0, 1e-5))
    yo = y[:, :, shard.own[0]:shard.own[1]].double()
    sums = torch.cat([yo.sum(dim=(0, 2, 3, 4)), (yo * yo).sum(dim=(0, 2, 3, 4))])
    shard.all_reduce(sums, "emu")
    c = y.shape[1]
    mean = sums[:c] / shard.m_global
    var = sums[c:] / shard.m_global - mean * mean
    bc = lambda v: v.float()[None, :, None, None, None]  # noqa: E731
    z = F.relu((y - bc(mean)) * bc(torch.rsqrt(var + 1e-5)) * w[None, :, None, None, None] + b[None, :, None, None, None])
    d = torch.arange(y.shape[2])
    mask = ((d >= shard.valid[0]) & (d < shard.valid[1])).float()[None, None, :, None, None]
    return z * mask


def _emu_mode_conv(x, gate_in, params, bn, training, conv_type="normal", precision=None, shard=None):
    k5, k3, k1, a3, a5, gw, gb = params
    p = {"expert_conv5x5_conv": k5, "expert_conv3x3_conv": k3, "expert_conv1x1_conv": k1, "expert_avg3x3_conv": a3,
         "expert_avg5x5_conv": a5}
    g = otc.gate_softmax(gw, gb, gate_in.long(), k5.shape[0])
    w = otc.reparam(p, "", g)
    y = torch.cat([F.conv3d(x[i:i + 1], w[i], padding=2) for i in range(x.shape[0])], dim=0)
    if conv_type == "normal":
        y = _bn_relu_spec(y, bn, shard)
    return y


def _emu_down(x, conv_w, bn, training, shard=None, precision=None):
    return _bn_relu_spec(F.conv3d(x, conv_w, stride=2), (bn.weight, bn.bias), shard)


def _emu_up(x, convt_w, bn, training, shard=None, precision=None):
    return _bn_relu_spec(F.conv_transpose3d(x, convt_w, stride=2), (bn.weight, bn.bias), shard)


def _net_rank(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from repmode_b200 import functional as Fm, sharded
        from repmode_b200.nn_modules import MoDEConv, Net
        Fm.mode_conv, Fm.down_conv_bn_relu, Fm.up_conv_bn_relu = _emu_mode_conv, _emu_down, _emu_up
        MoDEConv.forward = lambda self, x, t: _emu_mode_conv(       # replicated (gathered) levels call the modules
            x, t, self._params(), sharded._bn(self), self.training, self.conv_type)
        torch.manual_seed(1)
        net = Net(argparse.Namespace(adopted_datasets=list(range(6)), gpu_ids=-1), mult_chan=4).train()
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm3d):
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.uniform_(-0.3, 0.3)
        NB, D, H, W = 2, 32 * world, 16, 16
        g = torch.Generator().manual_seed(7)
        x = torch.randn(NB, 1, D, H, W, generator=g)
        t = torch.tensor([4, 1])
        dl = D // world
        with torch.no_grad():
            yl = sharded.sharded_net_forward(net, x[:, :, rank * dl:(rank + 1) * dl].contiguous(), t, D)
            p = {k: v.detach() for k, v in net.state_dict().items()}
            yo = otc.net_forward(p, x, t, True)
        ref = yo[:, :, rank * dl:(rank + 1) * dl]
        ret[rank] = (float((yl - ref).abs().max()), float(yo.abs().max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_net_forward_matches_unsharded_oracle(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_net_rank, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        err, scale = ret[r]
        assert err <= 2e-4 * max(scale, 1e-3), (r, err, scale)
