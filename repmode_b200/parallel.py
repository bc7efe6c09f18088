"""Multi-GPU plumbing for the MoDE-conv path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the
CPU tests).  The reference's only multi-GPU code is torch.nn.DataParallel (fnet/fnet_model.py:40-44, latently
broken for >1 GPU); nothing here mirrors it.

Two decompositions:
  * volume / batch sharding: independent volumes per rank, no data-path collective; the only exchange is
    `sync_gradients` (one flat all-reduce of the block's parameters per step).
  * D-axis sharding of one large volume: each rank owns a contiguous slab of d-planes; a two-conv stage
    (MoDESubNet2Conv, RepMode.py:111-120) needs ONE exchange of 4 boundary planes per side (`halo_exchange`),
    zero padding only at the GLOBAL faces, and BatchNorm statistics all-reduced over OWNED voxels
    (`allreduce_bn_sums`).  A single 2-plane exchange is wrong (1.5e-2), 4 planes are exact.
"""
import torch
import torch.distributed as dist


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def sync_gradients(params, group=None, average=False):
    """Data-parallel gradient exchange: flatten -> ONE all-reduce -> scatter back (in place)."""
    world, _ = _world(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat.div_(world)
    off = 0
    views = []
    for g in grads:
        views.append(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    torch._foreach_copy_(grads, views)


def slab_bounds(D, world, rank):
    """Contiguous d-plane slab [lo, hi) owned by `rank` (D must split evenly so stride-2 levels stay aligned)."""
    if D % world != 0:
        raise ValueError(f"D={D} is not divisible by the number of ranks {world}")
    n = D // world
    return rank * n, (rank + 1) * n


def halo_exchange(x_local, planes, group=None):
    """x_local: [N, D_local, H, W, C] (NDHWC slab).  Returns [N, D_local + 2*planes, H, W, C]: the slab with
    `planes` boundary planes of each D-neighbour attached; global faces get zeros (= the conv's zero padding).
    One grouped send/recv per call (batch_isend_irecv -> a single NCCL group)."""
    world, rank = _world(group)
    n, d, h, w, c = x_local.shape
    if planes > d and world > 1:
        raise ValueError(f"halo of {planes} planes exceeds the local slab depth {d}: re-shard or exchange per conv")
    lo = torch.zeros((n, planes, h, w, c), dtype=x_local.dtype, device=x_local.device)
    hi = torch.zeros_like(lo)
    if world > 1:
        ops = []
        send_lo = x_local[:, :planes].contiguous()
        send_hi = x_local[:, d - planes:].contiguous()
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, send_lo, rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, lo, rank - 1, group))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, send_hi, rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, hi, rank + 1, group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return torch.cat((lo, x_local, hi), dim=1)


def halo_reduce(g_ext, planes, group=None):
    """Adjoint of halo_exchange: g_ext [N, D_local + 2*planes, ...] -> [N, D_local, ...] with the gradient that
    landed in each halo sent back to its owner and accumulated (global-face halos are dropped)."""
    world, rank = _world(group)
    d = g_ext.shape[1] - 2 * planes
    out = g_ext[:, planes:planes + d].clone()
    if world > 1:
        ops = []
        send_lo = g_ext[:, :planes].contiguous()
        send_hi = g_ext[:, planes + d:].contiguous()
        recv_lo = torch.zeros_like(send_lo)
        recv_hi = torch.zeros_like(send_hi)
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, send_lo, rank - 1, group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, rank - 1, group))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, send_hi, rank + 1, group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, rank + 1, group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        out[:, :planes] += recv_lo          # what the lower neighbour computed on OUR first planes
        out[:, d - planes:] += recv_hi
    return out


class HaloExchange(torch.autograd.Function):
    """Differentiable halo exchange (forward: attach neighbour planes; backward: return + accumulate)."""

    @staticmethod
    def forward(ctx, x_local, planes, group):
        ctx.planes, ctx.group = planes, group
        return halo_exchange(x_local, planes, group)

    @staticmethod
    def backward(ctx, g_ext):
        return halo_reduce(g_ext.contiguous(), ctx.planes, ctx.group), None, None


def allreduce_bn_sums(sums, group=None):
    """Global BatchNorm statistics under D-sharding: all-reduce the per-channel [sum, sum of squares] (fp64, 2*C
    values) accumulated over OWNED voxels only."""
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(sums, group=group)
    return sums
