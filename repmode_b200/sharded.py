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
unsharded kernels (tools/check_sharded.py).
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
