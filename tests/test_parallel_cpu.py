"""CPU, gloo, world_size 2: host-side multi-rank logic of repmode_b200/parallel.py -- flat gradient all-reduce,
halo exchange and its adjoint, and the D-sharded two-conv stage algorithm (4-plane halo, zero padding only at
global faces, BatchNorm statistics over owned voxels) against the unsharded oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from repmode_b200 import parallel as par


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _sync_grads(rank, world):
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float(rank + 1 + i))
    par.sync_gradients(ps)
    return [p.grad.clone() for p in ps]


def test_sync_gradients():
    out = _spawn(_sync_grads)
    for r in range(2):
        assert torch.equal(out[r][0], torch.full((3, 4), 3.0))      # (0+1) + (1+1)
        assert torch.equal(out[r][1], torch.full((5,), 5.0))        # (0+2) + (1+2)


def _halo(rank, world):
    torch.manual_seed(1)
    full = torch.randn(1, 8, 3, 3, 2)
    lo, hi = par.slab_bounds(8, world, rank)
    x = full[:, lo:hi].clone().requires_grad_(True)
    ext = par.HaloExchange.apply(x, 2, None)
    # adjoint check: <ext, g> == <x, halo_reduce(g)> summed over ranks
    g = torch.randn(ext.shape, generator=torch.Generator().manual_seed(10 + rank))
    (ext * g).sum().backward()
    return ext.detach(), x.grad, g, full


def test_halo_exchange_and_adjoint():
    out = _spawn(_halo)
    full = out[0][3]
    padded = F.pad(full, [0, 0, 0, 0, 0, 0, 2, 2])                       # zero planes at the global faces
    for r in range(2):
        assert torch.equal(out[r][0], padded[:, 4 * r:4 * r + 8])
    # gradient of sum_r <ext_r, g_r> w.r.t. the full volume, computed directly
    gfull = torch.zeros_like(padded)
    for r in range(2):
        gfull[:, 4 * r:4 * r + 8] += out[r][2]
    for r in range(2):
        assert torch.allclose(out[r][1], gfull[:, 2 + 4 * r:2 + 4 * r + 4], atol=1e-6)


# ---------------------------------------------------------------------------------- D-sharded two-conv stage
def _bn_relu_global(y_ext, owned, gamma, beta, D_global, d_off):
    """train-mode BN over OWNED planes only (all-reduced), applied to all planes; planes outside the global
    volume are forced to zero afterwards (they are conv2's zero padding, not data)."""
    yo = y_ext[:, :, owned[0]:owned[1]]
    sums = torch.stack([yo.double().sum(dim=(0, 2, 3, 4)), (yo.double() ** 2).sum(dim=(0, 2, 3, 4))]).reshape(-1)
    cnt = torch.tensor([float(yo.numel() // yo.shape[1])], dtype=torch.float64)
    par.allreduce_bn_sums(sums)
    dist.all_reduce(cnt)
    c = y_ext.shape[1]
    mean = sums[:c] / cnt
    var = sums[c:] / cnt - mean ** 2
    b = lambda v: v.float()[None, :, None, None, None]  # noqa: E731
    z = (y_ext - b(mean)) * b(1.0 / torch.sqrt(var + 1e-5)) * gamma[None, :, None, None, None] + beta[None, :, None, None, None]
    z = F.relu(z)
    dglob = torch.arange(y_ext.shape[2]) + d_off
    mask = ((dglob >= 0) & (dglob < D_global)).float()[None, None, :, None, None]
    return z * mask


def _stage(rank, world):
    torch.manual_seed(3)
    N, C, D, H, W = 1, 3, 8, 6, 6
    x = torch.randn(N, C, D, H, W)
    w1 = torch.randn(4, C, 5, 5, 5) * 0.1
    w2 = torch.randn(4, 4, 5, 5, 5) * 0.1
    g1, b1, g2, b2 = torch.rand(4) + 0.5, torch.randn(4) * 0.1, torch.rand(4) + 0.5, torch.randn(4) * 0.1
    # unsharded oracle stage (conv -> BN(train) -> ReLU twice), every rank computes it for comparison
    def bnr(y, g, b):
        return F.relu(F.batch_norm(y, None, None, g, b, True, 0.0, 1e-5))
    ref = bnr(F.conv3d(bnr(F.conv3d(x, w1, padding=2), g1, b1), w2, padding=2), g2, b2)
    # sharded: ONE 4-plane exchange, conv1 on D_local+8 (valid on the inner D_local+4), conv2 valid on D_local
    lo, hi = par.slab_bounds(D, world, rank)
    dl = hi - lo
    xl = x[:, :, lo:hi].permute(0, 2, 3, 4, 1).contiguous()
    ext = par.halo_exchange(xl, 4).permute(0, 4, 1, 2, 3)                         # [N,C,dl+8,H,W], d offset lo-4
    y1 = F.conv3d(ext, w1, padding=2)[:, :, 2:-2]                                 # dl+4 planes, d offset lo-2
    a1 = _bn_relu_global(y1, (2, 2 + dl), g1, b1, D, lo - 2)
    y2 = F.conv3d(a1, w2, padding=2)[:, :, 2:-2]                                  # dl planes, d offset lo
    a2 = _bn_relu_global(y2, (0, dl), g2, b2, D, lo)
    return float((a2 - ref[:, :, lo:hi]).abs().max()), float(ref.abs().max())


def test_d_sharded_stage_matches_unsharded():
    for err, scale in _spawn(_stage):
        assert err <= 2e-5 * max(scale, 1.0), (err, scale)


# ---------------------------------------------------------------------------------- peer.TorchComm (exchange interface)
def _comm_steps(rank, world):
    from repmode_b200 import peer
    comm = peer.TorchComm()
    torch.manual_seed(5)
    full = torch.randn(1, 6 * world, 3, 4, 2)
    ext = comm.alloc("x", (1, 6 + 4, 3, 4, 2), torch.float32, "cpu")
    ext[0, 2:8] = full[0, 6 * rank:6 * rank + 6]
    comm.halo_fill(ext, 2, "t.x")
    v = torch.arange(5, dtype=torch.float64) + rank
    comm.all_reduce(v, "t.v")
    return ext.clone(), v, full, comm.n_collectives


def test_torch_comm_halo_fill_and_all_reduce():
    out = _spawn(_comm_steps)
    full = out[0][2]
    padded = F.pad(full, [0, 0, 0, 0, 0, 0, 2, 2])
    for r in range(2):
        assert torch.equal(out[r][0], padded[:, 6 * r:6 * r + 10])
        assert torch.equal(out[r][1], 2 * torch.arange(5, dtype=torch.float64) + 1)
        assert out[r][3] == 2
