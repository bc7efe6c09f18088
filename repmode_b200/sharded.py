"""D-axis sharding of one large volume across the GPUs of a box: a two-conv stage
(MoDESubNet2Conv, reference fnet/nn_modules/RepMode.py:111-120) on a slab of d-planes with ONE halo exchange.

Per stage and rank (slab [lo, lo+dl) of a D-plane volume):
  1. attach 4 planes from each D-neighbour (zeros at the global faces)            -> dl + 8 planes   [NCCL send/recv]
  2. conv1 + BN + ReLU on the extended slab; BatchNorm sums over OWNED planes, all-reduced; planes beyond the
     global volume are forced to zero afterwards (they are conv2's zero padding)  -> keep the inner dl + 4 planes
  3. conv2 + BN + ReLU; keep the dl owned planes.
Backward is the autograd mirror: dgrad/wgrad on the extended slabs, BN-backward sums all-reduced with the mean
terms applied on owned planes only, halo gradients sent back to their owners and accumulated; parameter
gradients are partial per rank and summed by the usual data-parallel all-reduce (parallel.sync_gradients).
Verified on CPU/gloo against the unsharded oracle (tests/test_parallel_cpu.py) and on 2 GPUs against the
unsharded kernels (tests/check_sharded.py).
"""
import torch
import torch.distributed as dist

from . import functional as Fm
from . import parallel as par

HALO = 4     # two stacked 5^3 convs reach 4 planes; a single 2-plane exchange is wrong by 1.5e-2


def _bn(mod):
    if mod.conv_type != "normal":
        return None
    m = mod.subsequent_layer[0]
    return (m.weight, m.bias, m.running_mean, m.running_var)


def sharded_stage(stage, x_local, t, d_global, group=None):
    """stage: MoDESubNet2Conv; x_local: this rank's slab [N, C, dl, H, W] of an [N, C, d_global, H, W] volume.
    Returns this rank's slab of the stage output."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n, _, dl, h, w = x_local.shape
    lo = rank * dl
    if dl * world != d_global:
        raise ValueError("slabs must tile the volume evenly")
    m_global = n * d_global * h * w
    xe = par.HaloExchange.apply(Fm.to_ndhwc(x_local), HALO, group)               # [N, dl+8, H, W, C], plane 0 = lo-4
    xe = Fm.from_ndhwc(xe)
    c1, c2 = stage.conv1, stage.conv2
    for c in (c1, c2):
        if c.training and c.conv_type == "normal":
            c.subsequent_layer[0].num_batches_tracked.add_(1)
    s1 = Fm.ShardSpec(own=(HALO, HALO + dl), valid=(max(0, HALO - lo), min(dl + 2 * HALO, d_global - lo + HALO)),
                      m_global=m_global, group=group)
    a1 = Fm.mode_conv(xe, t, c1._params(), _bn(c1), c1.training, c1.conv_type, c1.precision, s1)
    a1 = a1[:, :, 2:dl + 6]                                                        # conv1 is valid here; plane 0 = lo-2
    s2 = Fm.ShardSpec(own=(2, 2 + dl), valid=(max(0, 2 - lo), min(dl + 4, d_global - lo + 2)), m_global=m_global,
                      group=group)
    a2 = Fm.mode_conv(a1, t, c2._params(), _bn(c2), c2.training, c2.conv_type, c2.precision, s2)
    return a2[:, :, 2:2 + dl]


# ---------------------------------------------------------------------------------------------- whole U-Net
class _AllGatherD(torch.autograd.Function):
    """[N,C,dl,H,W] slabs -> the full [N,C,dl*world,H,W] volume on every rank.  Downstream of it the computation is
    replicated but each rank's loss only covers its own slab, so the gradients arriving here are PARTIAL sums:
    backward = sum over ranks, then keep the own slab (reduce-scatter)."""

    @staticmethod
    def forward(ctx, x_local, group):
        ctx.group = group
        world = dist.get_world_size(group)
        parts = [torch.empty_like(x_local) for _ in range(world)]
        dist.all_gather(parts, x_local.contiguous(), group=group)
        return torch.cat(parts, dim=2)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, group=ctx.group)
        world, rank = dist.get_world_size(ctx.group), dist.get_rank(ctx.group)
        dl = g.shape[2] // world
        return g[:, :, rank * dl:(rank + 1) * dl].contiguous(), None


def _full_spec(n, d_local, h, w, d_global, group):
    """ShardSpec of a halo-free slab tensor (stride-2 levels): every local plane is owned and valid."""
    return Fm.ShardSpec(own=(0, d_local), valid=(0, d_local), m_global=n * d_global * h * w, group=group)


def sharded_net_forward(net, x_local, t, d_global, group=None, probe=None, replicated=False):
    """Net.forward (reference RepMode.py:51-71) on a D-slab of one volume.  Levels whose local depth is at least the
    halo (4 planes) run sharded (one halo exchange per two-conv stage, global BatchNorm statistics); deeper, tiny
    levels are gathered and computed redundantly on every rank (their tensors are a few MB), then re-sliced on the
    way up.  x_local: [N,1,dl,H,W]; returns this rank's [N,1,dl,H,W] slab of the prediction."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1 and probe is None:
        return net(x_local, t)
    t = t.to(device=x_local.device, dtype=torch.int32).reshape(-1)

    def keep(name, v):              # diagnostics: remember (and keep the gradient of) named intermediates
        if probe is not None:
            v.retain_grad()
            probe[name] = v
        return v
    enc = [net.encoder_block1, net.encoder_block2, net.encoder_block3, net.encoder_block4]
    dec = [net.decoder_block4, net.decoder_block3, net.decoder_block2, net.decoder_block1]
    n = x_local.shape[0]
    training = net.training

    def bump(bn):
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)

    x = x_local
    dg = d_global                       # global depth at the current level
    sharded = not (replicated or world == 1)   # is x a slab (True) or the replicated full tensor (False)?
    skips = []
    for li, blk in enumerate(enc):
        if sharded and dg // world < HALO:                      # too thin for a 4-plane halo: replicate from here down
            x = _AllGatherD.apply(x, group)
            sharded = False
        if sharded:
            x_skip = sharded_stage(blk.conv_more, x, t, dg, group)
        else:
            x_skip = blk.conv_more(x, t)
        keep(f"enc{li + 1}.skip", x_skip)
        skips.append((x_skip, sharded))
        bn = blk.conv_down[1]
        bump(bn)
        _, _, dl, h, w = x_skip.shape
        spec = _full_spec(n, dl // 2, h // 2, w // 2, dg // 2, group) if sharded else None
        x = keep(f"enc{li + 1}.down", Fm.down_conv_bn_relu(x_skip, blk.conv_down[0].weight, bn, training, spec))
        dg //= 2
    if sharded and dg // world < HALO:
        x = _AllGatherD.apply(x, group)
        sharded = False
    x = keep("bottle", sharded_stage(net.bottle_block, x, t, dg, group) if sharded else net.bottle_block(x, t))
    for li, blk in enumerate(dec):
        x_skip, skip_sharded = skips.pop()
        bn = blk.convt[1]
        bump(bn)
        if skip_sharded and not sharded:                        # back to slabs: keep this rank's planes of the replica
            dl = x.shape[2] // world
            x = x[:, :, rank * dl:(rank + 1) * dl]
            sharded = True
        _, _, dl, h, w = x.shape
        spec = _full_spec(n, 2 * dl, 2 * h, 2 * w, 2 * dg, group) if sharded else None
        x = keep(f"dec{4 - li}.up", Fm.up_conv_bn_relu(x, blk.convt[0].weight, bn, training, spec))
        dg *= 2
        xc = torch.cat((x_skip, x), 1)
        x = keep(f"dec{4 - li}.out", sharded_stage(blk.conv_less, xc, t, dg, group) if sharded else blk.conv_less(xc, t))
    # conv_out: a single 5^3 conv, no BatchNorm -> 2-plane halo
    c = net.conv_out
    if sharded:
        dl = x.shape[2]
        xe = Fm.from_ndhwc(par.HaloExchange.apply(Fm.to_ndhwc(x), 2, group))
        y = Fm.mode_conv(xe, t, c._params(), None, c.training, c.conv_type, c.precision)
        return y[:, :, 2:2 + dl]
    y = c(x, t)
    if replicated or world == 1:
        return y
    dl = y.shape[2] // world
    return y[:, :, rank * dl:(rank + 1) * dl]
