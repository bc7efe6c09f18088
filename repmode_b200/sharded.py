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
import os

import torch
import torch.distributed as dist

from . import functional as Fm
from . import parallel as par

HALO = 4     # two stacked 5^3 convs reach 4 planes; a single 2-plane exchange is wrong by 1.5e-2


def _bn(mod):
    if mod.conv_type != "normal":
        return None
    m = mod.subsequent_layer[0]
    return (m.weight, m.bias, m.running_mean, m.running_var, m.eps, m.momentum)


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


# ------------------------------------------------------------------------------------- one MoDEConv on a D-slab
BLOCK_HALO = 2      # a single 5^3 conv reaches 2 planes
# exchange steps fused into the producing / consuming kernels when the comm supports it (peer.PeerComm); 0 = separate
# put / wait / sum launches (the A/B arm)
FUSED_EXCHANGE = os.environ.get("REPMODE_FUSED_EXCHANGE", "1") == "1"


class ShardedConvFunction(torch.autograd.Function):
    """One MoDEConv (conv_type 'normal', train mode: RepMode.py:194-214) on this rank's D-slab of ONE volume -- the headline
    block of BASELINE.json under the D-axis split.  Per step and rank:

      forward   x slab -> conv operand written into the INTERIOR of a haloed buffer [1, D+4, H, W, Ci]; `comm.halo_fill`
                brings the 2 boundary planes of each neighbour (zeros at the global faces); K2 computes ONLY the D owned
                output planes from the haloed operand (mode_conv3d_ex: no redundant planes); BatchNorm sums (all output
                planes are owned) are all-reduced; finalize + apply as on one GPU.
      backward  BatchNorm-backward sums all-reduced; dy written into the interior of a second haloed buffer and exchanged
                the same way; K4 runs on (haloed x) x (owned dy), K3 on the haloed dy -> dx of the owned planes; K1b; the
                partial parameter gradients of the ranks are summed by ONE all-reduce of the flat gradient buffer, so the
                gradients this node returns are already the sums over the whole volume (do NOT all-reduce them again).

    5 exchange steps per training step (2 halo, 2 BatchNorm, 1 gradient), all through `comm` (peer.PeerComm: stores into the
    neighbour's memory over NVLink; peer.TorchComm: NCCL / gloo).  N = 1 per rank (one volume)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    @Fm._on_device_of_first
    def forward(ctx, x, gate_in, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, running_mean, running_var, eps, momentum,
                precision, comm, d_global, tag):
        import ctypes
        from . import lib as _lib
        Fm._require_cuda(x, gate_in, k5)
        lib = _lib.load()
        n, ci_x, d, h, wd = x.shape
        if n != 1:
            raise RuntimeError("sharded MoDEConv: one volume per rank (N = 1)")
        layer, ci, co = Fm._layer(k5, k3, k1, a3, a5, gate_w, gate_b)
        if ci_x != ci:
            raise RuntimeError(f"MoDEConv: input has {ci_x} channels, layer expects {ci}")
        dev = x.device
        gate_in = gate_in.contiguous().float() if gate_in.dtype.is_floating_point else gate_in.to(torch.int32).contiguous()
        use_umma = precision == "f16" and Fm.umma_shape_ok(ci, co, d, h, wd) and ci % 32 == 0 and co % 32 == 0
        dtype = _lib.MODE_F16 if use_umma else _lib.MODE_F32
        tdt = torch.float16 if use_umma else torch.float32
        H2 = BLOCK_HALO
        sample_u = Fm._sample_index(1, dev, False)
        w_scale = Fm.W_SCALE_F16 if use_umma else 1.0
        needs_dx = ctx.needs_input_grad[0]

        xn = Fm.to_ndhwc(x)
        x_ext = comm.alloc((tag, "x_ext", str(tdt)), (1, d + 2 * H2, h, wd, ci), tdt, dev)
        k1_fork = Fm._Fork(dev, True)
        sums = torch.empty(2 * co, dtype=torch.float64, device=dev)      # cleared on the side stream, next to K1
        g, w_fwd, w_dg = Fm.reparam_fwd(layer, gate_in[:1].contiguous(), 1, ci, co, dtype, needs_dx, w_scale, fork=k1_fork,
                                        post=sums.zero_)
        fused = FUSED_EXCHANGE and getattr(comm, "fused", False) and comm.world > 1
        if use_umma and fused:
            # the cast kernel stores its boundary planes straight into the neighbours' halo planes and signals
            hp, wait_x = comm.halo_push_desc(x_ext, H2, tag + ".x")
            _lib.check(lib.mode_cast_f16_ex(Fm._p(xn), Fm._p(x_ext[0, H2]), xn.numel(), 1.0, None, ctypes.byref(hp),
                                            Fm._stream()), "mode_cast_f16")
            comm.halo_wait(wait_x)
        else:
            if use_umma:
                _lib.check(lib.mode_cast_f16(Fm._p(xn), Fm._p(x_ext[0, H2]), xn.numel(), 1.0, None, Fm._stream()),
                           "mode_cast_f16")
            else:
                x_ext[0, H2:H2 + d].copy_(xn[0])
            comm.halo_fill(x_ext, H2, tag + ".x")
        k1_fork.join()

        m_rows = d * h * wd
        m_global = d_global * h * wd
        mean = torch.empty(co, dtype=torch.float32, device=dev)
        invstd = torch.empty(co, dtype=torch.float32, device=dev)
        scale = torch.empty(co, dtype=torch.float32, device=dev)
        shift = torch.empty(co, dtype=torch.float32, device=dev)
        if fused:
            # K2's last CTA broadcasts the slab's sums; the finalize kernel waits for every rank's and adds them up
            pp, pg = comm.reduce_desc(2 * co, tag + ".bnf")
            y = Fm.conv3d(x_ext, dtype, w_fwd, sample_u, 1, d, h, wd, ci, co, None, sums, out_scale=1.0 / w_scale,
                          halo=(d + 2 * H2, H2), stats_push=pp)
            _lib.check(lib.mode_bn_finalize_ex(Fm._p(sums), m_global, co, Fm._p(bn_w), Fm._p(bn_b), float(eps),
                                               float(momentum), Fm._p(mean), Fm._p(invstd), Fm._p(scale), Fm._p(shift),
                                               Fm._p(running_mean), Fm._p(running_var), ctypes.byref(pg), Fm._stream()),
                       "mode_bn_finalize")
        else:
            y = Fm.conv3d(x_ext, dtype, w_fwd, sample_u, 1, d, h, wd, ci, co, None, sums, out_scale=1.0 / w_scale,
                          halo=(d + 2 * H2, H2))
            comm.all_reduce(sums, tag + ".bnf")
            _lib.check(lib.mode_bn_finalize(Fm._p(sums), m_global, co, Fm._p(bn_w), Fm._p(bn_b), float(eps),
                                            float(momentum), Fm._p(mean), Fm._p(invstd), Fm._p(scale), Fm._p(shift),
                                            Fm._p(running_mean), Fm._p(running_var), Fm._stream()), "mode_bn_finalize")
        out = torch.empty_like(y)
        _lib.check(lib.mode_bn_apply_relu(Fm._p(y), m_rows, co, Fm._p(scale), Fm._p(shift), 1, Fm._p(out), None, 1.0, None,
                                          Fm._stream()), "mode_bn_apply_relu")
        ctx.save_for_backward(y, g, w_dg, gate_in, sample_u, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, mean, invstd)
        ctx.x_ext = x_ext                 # persistent exchange buffer: not a saved tensor (it is rewritten every step)
        ctx.cfg = (d, h, wd, ci, co, use_umma, needs_dx, m_global, comm, tag, fused)
        return Fm.from_ndhwc(out)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @Fm._on_device_of_first
    def backward(ctx, dout):
        import ctypes
        from . import lib as _lib
        lib = _lib.load()
        y, g, w_dg, gate_in, sample_u, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, mean, invstd = ctx.saved_tensors
        d, h, wd, ci, co, use_umma, needs_dx, m_global, comm, tag, fused = ctx.cfg
        x_ext = ctx.x_ext
        dev = dout.device
        dtype = _lib.MODE_F16 if use_umma else _lib.MODE_F32
        tdt = torch.float16 if use_umma else torch.float32
        H2 = BLOCK_HALO
        doutn = Fm.to_ndhwc(dout)
        m_rows = d * h * wd
        ws = torch.empty(int(lib.mode_bn_bwd_workspace_bytes(co)), dtype=torch.uint8, device=dev)
        planes = _lib.ModePlanes(h * wd, d, 0, d, 0, d, m_global)
        pl = ctypes.byref(planes)
        dy_ext = comm.alloc((tag, "dy_ext", str(tdt)), (1, d + 2 * H2, h, wd, co), tdt, dev)
        dgamma = torch.empty(co, dtype=torch.float32, device=dev)
        dbeta = torch.empty(co, dtype=torch.float32, device=dev)
        dy_s2 = torch.empty(2, dtype=torch.float32, device=dev) if use_umma else None
        dy_int = dy_ext[0, H2:H2 + d]
        apply_args = (Fm._p(y), Fm._p(doutn), m_rows, co, Fm._p(bn_w), Fm._p(bn_b), Fm._p(mean), Fm._p(invstd),
                      Fm._p(dgamma), Fm._p(dbeta), None if use_umma else Fm._p(dy_int), Fm._p(dy_int) if use_umma else None,
                      Fm._p(dy_s2), pl, Fm._p(ws))
        # {sum dz, sum dz*xhat, max|dz|, max|xhat|}[co] travel as ONE fp64 vector: the fp16 scale of dy must be the SAME on
        # every rank (halo planes travel in fp16) and the sum of the ranks' maxima bounds the global maximum
        if fused:
            # the reduce kernel's last block broadcasts the vector; the scale kernel in front of the apply pass gathers it;
            # the apply pass stores the boundary planes of dy straight into the neighbours' halo planes
            pp, pg = comm.reduce_desc(4 * co, tag + ".bnb")
            _lib.check(lib.mode_bn_relu_bwd_reduce_ex(Fm._p(y), Fm._p(doutn), m_rows, co, Fm._p(bn_w), Fm._p(bn_b),
                                                      Fm._p(mean), Fm._p(invstd), pl, Fm._p(ws), ctypes.byref(pp),
                                                      Fm._stream()), "mode_bn_relu_bwd_reduce")
            hp, wait_dy = comm.halo_push_desc(dy_ext, H2, tag + ".dy")
            _lib.check(lib.mode_bn_relu_bwd_apply_ex(*apply_args, ctypes.byref(pg), ctypes.byref(hp), Fm._stream()),
                       "mode_bn_relu_bwd_apply")
            # the wait for the neighbours' dy planes is issued in front of K3, the only consumer of the dy halo: K4 reads the
            # OWNED dy planes only, so the exchange's flight time and the rank skew hide behind it (r2v)
        else:
            _lib.check(lib.mode_bn_relu_bwd_reduce(Fm._p(y), Fm._p(doutn), m_rows, co, Fm._p(bn_w), Fm._p(bn_b), Fm._p(mean),
                                                   Fm._p(invstd), pl, Fm._p(ws), Fm._stream()), "mode_bn_relu_bwd_reduce")
            comm.all_reduce(ws.view(torch.float64), tag + ".bnb")
            _lib.check(lib.mode_bn_relu_bwd_apply(*apply_args, Fm._stream()), "mode_bn_relu_bwd_apply")
            comm.halo_fill(dy_ext, H2, tag + ".dy")
        inv = dy_s2[1:2] if use_umma else None
        d_weff, finish_wgrad, _ = Fm.conv3d_wgrad(x_ext, dy_int, dtype, 1, d, h, wd, ci, co, inv, halo=(d + 2 * H2, H2),
                                                  two_phase=True)
        dx = None
        dg_fork = Fm._Fork(dev, needs_dx)
        if fused and not needs_dx:
            comm.halo_wait(wait_dy)                     # nobody reads the halo, but the counter protocol still advances
        if needs_dx:
            dxn = torch.empty((1, d, h, wd, ci), dtype=torch.float32, device=dev)
            with dg_fork:
                if fused:
                    comm.halo_wait(wait_dy)
                Fm.conv3d(dy_ext, dtype, w_dg, sample_u, 1, d, h, wd, co, ci, inv, None,
                          out_scale=(1.0 / Fm.W_SCALE_F16) if use_umma else 1.0, out=dxn, halo=(d + 2 * H2, H2))
        finish_wgrad()                                  # K4's slab reduce, now next to K3 instead of in front of it
        layer, _, _ = Fm._layer(k5, k3, k1, a3, a5, gate_w, gate_b)
        params = (k5, k3, k1, a3, a5, gate_w, gate_b)
        sizes = [p.numel() for p in params]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        outs, off = [], 0
        for p, sz in zip(params, sizes):
            outs.append(flat[off:off + sz].view_as(p))
            off += sz
        wsb = torch.empty(max(int(lib.mode_reparam_bwd_workspace_bytes(ci, co, 1)), 16), dtype=torch.uint8, device=dev)
        ids, dense = (gate_in[:1], None) if not gate_in.dtype.is_floating_point else (None, gate_in[:1])
        _lib.check(lib.mode_reparam_bwd(ctypes.byref(layer), Fm._p(ids), Fm._p(dense), 1, Fm._p(sample_u), 1, Fm._p(g),
                                        Fm._p(d_weff), *[Fm._p(o) for o in outs], Fm._p(wsb), Fm._stream()),
                   "mode_reparam_bwd")
        # gradient sum over the slabs NEXT TO dgrad (side stream).  No hazard on the dy halo: a neighbour's next dy push comes
        # after ITS BatchNorm-backward all-reduce of the next step, which needs OUR contribution -- stream-ordered after the join
        comm.all_reduce(flat, tag + ".grad")
        dg_fork.join()
        if needs_dx:
            dx = Fm.from_ndhwc(dxn)
        return (dx, None, *outs, dgamma, dbeta, None, None, None, None, None, None, None, None)


def sharded_mode_conv(mod, x_local, t, comm, d_global, tag="blk"):
    """mod: a MoDEConv (conv_type 'normal') in train mode; x_local: this rank's slab [1, Ci, D_local, H, W] of one
    [1, Ci, d_global, H, W] volume.  Returns this rank's slab of the block output; after backward every parameter's .grad
    holds the gradient summed over all slabs."""
    if mod.conv_type != "normal" or not mod.training:
        raise RuntimeError("sharded_mode_conv: train-mode 'normal' MoDEConv only")
    bn = mod.subsequent_layer[0]
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return ShardedConvFunction.apply(x_local, t, *mod._params(), bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                     bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                                     mod.precision or Fm.default_precision(), comm, d_global, tag)


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
        x = keep(f"enc{li + 1}.down", Fm.down_conv_bn_relu(x_skip, blk.conv_down[0].weight, bn, training, spec,
                                                            precision=blk.conv_more.conv2.precision))
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
        x = keep(f"dec{4 - li}.up", Fm.up_conv_bn_relu(x, blk.convt[0].weight, bn, training, spec,
                                                         precision=blk.conv_less.conv1.precision))
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
