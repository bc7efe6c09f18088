"""Host side of the MoDE-conv hot path: torch.autograd.Function + thin wrappers that hand raw device
pointers and the current CUDA stream to the C ABI (include/repmode_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic step of
MoDEConv.forward / backward (fnet/nn_modules/RepMode.py:194-214 in the reference tree) runs in the
hand-written sm_100a kernels.  There is no CPU / eager fallback: a non-CUDA tensor raises.
"""
import ctypes
import functools
import os
import weakref

import torch

from . import lib as _lib

E = 5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def default_precision():
    """'f16' = tcgen05 tensor-core path (fp16 operands with power-of-two scaling, fp32 accumulate; same
    10-bit mantissa as TF32) wherever the layer shape allows, 'f32' = SIMT fp32 everywhere."""
    return os.environ.get("REPMODE_PRECISION", "f16")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device_of_first(fn):
    """Run an autograd forward / backward with the CUDA device of its first tensor argument made current: the C library
    launches on the CURRENT device and stream (tensor maps, SM count, error flag are per device), while the reference only
    ever moves tensors (`.to(cuda:gpu_ids[0])`, fnet_model.py:53) and never calls set_device -- so Model(opts, gpu_ids=1)
    in a process whose current device is 0 must still run on cuda:1.  CPU tensors pass through to the CUDA-only check."""
    @functools.wraps(fn)
    def wrapped(ctx, first, *args, **kwargs):
        if torch.is_tensor(first) and first.is_cuda:
            with torch.cuda.device(first.device):
                return fn(ctx, first, *args, **kwargs)
        return fn(ctx, first, *args, **kwargs)
    return wrapped


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("repmode_b200: the MoDE-conv path is CUDA-only (sm_100a); got a CPU tensor. "
                               "There is no CPU fallback by design.")


OVERLAP = os.environ.get("REPMODE_OVERLAP", "1") == "1"    # run the independent branches of a step on a second stream
_side_streams = {}


def _side_stream(dev):
    """One extra CUDA stream per device for work that is independent of the main chain (K1 next to the operand cast,
    dgrad next to K1b).  Tensors are always ALLOCATED on the caller's stream before the fork and joined back into it, so
    the caching allocator never sees a cross-stream hand-off; under CUDA-graph capture the fork/join become parallel
    graph branches."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=dev)
    return st


class _Fork:
    """`with _Fork(dev, enabled):` -- launches inside run on the side stream, ordered after everything already enqueued
    on the caller's stream; .join() makes the caller's stream wait for them."""

    def __init__(self, dev, enabled=True):
        self.enabled = enabled and OVERLAP
        if self.enabled:
            self.main = torch.cuda.current_stream(dev)
            self.side = _side_stream(dev)
            self.ctx = torch.cuda.stream(self.side)

    def __enter__(self):
        if self.enabled:
            self.side.wait_stream(self.main)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.enabled:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.enabled:
            self.main.wait_stream(self.side)


def to_ndhwc(x):
    """[N,C,D,H,W] (any strides) -> dense [N,D,H,W,C] fp32. Free when x is already channels_last_3d."""
    return x.permute(0, 2, 3, 4, 1).contiguous().float()


def from_ndhwc(y):
    """dense [N,D,H,W,C] -> logical [N,C,D,H,W] view (channels_last_3d strides, no copy)."""
    return y.permute(0, 4, 1, 2, 3)


def _layer(k5, k3, k1, a3, a5, gate_w, gate_b):
    co, ci = k5.shape[0], k5.shape[1]
    for t in (k5, k3, k1, a3, a5, gate_w, gate_b):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("repmode_b200: MoDEConv parameters must be contiguous fp32")
    L = _lib.ModeLayer(k5.data_ptr(), k3.data_ptr(), k1.data_ptr(), a3.data_ptr(), a5.data_ptr(), gate_w.data_ptr(),
                       gate_b.data_ptr(), ci, co, gate_w.shape[1])
    return L, ci, co


UMMA_WGRAD = os.environ.get("REPMODE_UMMA_WGRAD", "1") == "1"   # K4 on tcgen05 (wgrad_deep.cu); 0 -> SIMT fp32 wgrad


def _pad32(c):
    return (c + 31) // 32 * 32


def umma_shape_ok(ci, co, d, h, w):
    """Shapes the tcgen05 kernels take for forward (K=ci, N=co), dgrad (K=co, N=ci) and wgrad; channel counts that
    are not multiples of 32 (the U-Net stem Ci=1 and head Co=1) are zero-padded to 32 by the host side, so only
    W % 8 and the 128-channel pass granularity remain.  Everything else runs the SIMT fp32 kernels."""
    return w % 8 == 0 and os.environ.get("REPMODE_DISABLE_UMMA", "0") != "1"


def pad_channels(t, c_pad):
    """[..., C] -> [..., c_pad] with zero channels appended (no-op when already c_pad wide)."""
    if t.shape[-1] == c_pad:
        return t
    out = torch.zeros(t.shape[:-1] + (c_pad,), dtype=t.dtype, device=t.device)
    out[..., :t.shape[-1]] = t
    return out


W_SCALE_F16 = 256.0     # fixed power-of-two scale of the fp16 weight pack: |W_eff| <= max|expert| is O(1e-2..1) for
                        # BatchNorm-ed conv nets, so values stay far from fp16's 65504 (K1 saturates) and at least
                        # 2^-8 * 6e-5 = 2.4e-7 in absolute resolution; no per-call amax pass over the experts


def reparam_fwd(layer, gate_in, U, ci, co, dtype, want_dgrad, w_scale=1.0, fork=None, post=None):
    """K1. Returns g [U,5,Co], w_fwd, w_dgrad (packed, see header).  With `fork` (a _Fork) the kernels are launched
    on the side stream (outputs are still allocated on the caller's stream); the caller joins before using them.
    `post()` runs inside the fork after the K1 launches (small independent work kept off the caller's stream)."""
    lib = _lib.load()
    dev = gate_in.device
    tdt = torch.float16 if dtype == _lib.MODE_F16 else torch.float32
    g = torch.empty((U, E, co), dtype=torch.float32, device=dev)
    if dtype == _lib.MODE_F16:      # rows padded to 32; K1 leaves pad rows untouched -> zero them here when present
        mk = torch.zeros if (ci % 32 or co % 32) else torch.empty
        w_fwd = mk(U * lib.mode_packed_weight_elems_f16(ci, co), dtype=tdt, device=dev)
        w_dg = mk(U * lib.mode_packed_weight_elems_f16(co, ci), dtype=tdt, device=dev) if want_dgrad else None
    else:
        w_fwd = torch.empty(U * lib.mode_packed_weight_elems(ci, co), dtype=tdt, device=dev)
        w_dg = torch.empty(U * lib.mode_packed_weight_elems(co, ci), dtype=tdt, device=dev) if want_dgrad else None
    ids, dense = (gate_in, None) if not gate_in.dtype.is_floating_point else (None, gate_in)

    def launch():
        _lib.check(lib.mode_reparam_fwd(ctypes.byref(layer), _p(ids), _p(dense), U, _p(g), _p(w_fwd), _p(w_dg), dtype,
                                        float(w_scale), None, _stream()), "mode_reparam_fwd")
    if fork is not None:
        with fork:
            launch()
            if post is not None:
                post()
    else:
        launch()
        if post is not None:
            post()
    return g, w_fwd, w_dg


K1_GROUPED = os.environ.get("REPMODE_K1_GROUPED", "1") == "1"
_plan_streams = {}


def _plan_stream(dev):
    """A second side stream for the step-level K1 plan: its kernels must not sit in front of the per-layer forks, which
    join the ordinary side stream back into the caller's stream a few microseconds after they fork."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    st = _plan_streams.get(key)
    if st is None:
        st = _plan_streams[key] = torch.cuda.Stream(device=dev)
    return st


class K1Plan:
    """K1 of every row-eligible MoDEConv of a training step as grouped launches (mode_reparam_fwd_grouped: one K1 + one
    dgrad-pack launch per group) at the START of the step, on their own stream -- SURVEY.md section 8d's "single grouped
    launch, not 19x".  All layers of a step share the gate inputs (RepMode.py:51-71).  Two groups: the layers the forward
    needs first (a few narrow ones, ~10 us) and the rest (the wide, HBM-bound ones), which then runs UNDER the first conv
    layers instead of in front of its own.  `take(mod)` hands a layer its (g, w_fwd, w_dgrad) after making the caller's
    stream wait for the group's event."""

    def __init__(self):
        self.entries = {}

    @staticmethod
    def eligible(mod, x_is_cuda, w_last):
        ci, co = mod.in_chan, mod.out_chan
        return (x_is_cuda and (mod.precision or default_precision()) == "f16" and ci % 32 == 0 and co % 32 == 0
                and w_last % 8 == 0 and os.environ.get("REPMODE_DISABLE_UMMA", "0") != "1")

    @classmethod
    def build(cls, mods, gate_in, dev, first_group=3):
        """mods: eligible MoDEConv modules in execution order; gate_in: int32 task ids [U] (contiguous, on dev)."""
        lib = _lib.load()
        U = gate_in.shape[0]
        if not mods or U > 32 or len(mods) > 2 * _lib.REPARAM_GROUP_MAX:
            return None
        plan = cls()
        main = torch.cuda.current_stream(dev)
        side = _plan_stream(dev)
        items = []
        for m in mods:
            layer, ci, co = _layer(*m._params())
            g = torch.empty((U, E, co), dtype=torch.float32, device=dev)
            w_fwd = torch.empty(U * lib.mode_packed_weight_elems_f16(ci, co), dtype=torch.float16, device=dev)
            w_dg = torch.empty(U * lib.mode_packed_weight_elems_f16(co, ci), dtype=torch.float16, device=dev)
            items.append((m, layer, g, w_fwd, w_dg))
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for grp in (items[:first_group], items[first_group:]):
                for k0 in range(0, len(grp), _lib.REPARAM_GROUP_MAX):
                    part = grp[k0:k0 + _lib.REPARAM_GROUP_MAX]
                    arr = (_lib.ModeReparamItem * len(part))()
                    for a, (m, layer, g, w_fwd, w_dg) in zip(arr, part):
                        a.layer, a.g_out, a.w_fwd, a.w_dgrad = layer, g.data_ptr(), w_fwd.data_ptr(), w_dg.data_ptr()
                    _lib.check(lib.mode_reparam_fwd_grouped(arr, len(part), _p(gate_in), None, U, _lib.MODE_F16,
                                                            float(W_SCALE_F16), _stream()), "mode_reparam_fwd_grouped")
                if grp:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    for m, layer, g, w_fwd, w_dg in grp:
                        plan.entries[id(m)] = (g, w_fwd, w_dg, ev)
        return plan

    def take(self, mod, dev):
        ent = self.entries.pop(id(mod), None)
        if ent is None:
            return None
        g, w_fwd, w_dg, ev = ent
        torch.cuda.current_stream(dev).wait_event(ev)
        return g, w_fwd, w_dg

    def finish(self, dev):
        """Join whatever was not consumed (an exception, a skipped layer): the plan stream must be back in the caller's
        stream before a CUDA-graph capture can end."""
        if self.entries:
            torch.cuda.current_stream(dev).wait_stream(_plan_stream(dev))
            self.entries.clear()


def conv3d(x, dtype, w, sample_u, n, d, h, wd, k, nout, out_scale_dev=None, bn_sums=None, impl=0, stat_range=None,
           out_scale=1.0, out=None, halo=None, ep=None, y16=None, want_y=True, stats_push=None):
    """K2 / K3 through mode_conv3d_ex.  d = OUTPUT planes.
    halo = (Dx, x_off): x holds Dx >= d planes and output plane q is centred on input plane q + x_off (D-sharded slabs).
    ep = (scale[nout] | None, shift[nout] | None, relu): per-channel affine + ReLU fused into the epilogue (eval BatchNorm).
    y16 = (buffer fp16 [n, Dy16, h, wd, nout], y16_off, scale): fp16 copy of the result written at plane y16_off + q.
    want_y=False skips the fp32 result (returns None).
    stats_push = lib.ModePeerPush: the last CTA broadcasts bn_sums to every rank of a D-sharded volume (peer.PeerComm)."""
    lib = _lib.load()
    y = out
    if y is None and want_y:
        y = torch.empty((n, d, h, wd, nout), dtype=torch.float32, device=x.device)
    lo, hi = stat_range if stat_range is not None else (0, d)
    opts = None
    # deep small-volume layers run split along K and need scratch for the partial results (0 bytes for everything else)
    ws_bytes = int(lib.mode_conv3d_workspace_bytes(n, d, h, wd, k, nout, dtype)) if (impl in (0, 2, 3) and stats_push is None) else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
    if halo is not None or ep is not None or y16 is not None or stats_push is not None or ws is not None:
        o = _lib.ModeConvOpts()
        if ws is not None:
            o.splitk_ws, o.splitk_ws_bytes = ws.data_ptr(), ws_bytes
        if stats_push is not None:
            o.stats_push = ctypes.pointer(stats_push)
        if halo is not None:
            o.Dx, o.x_off = int(halo[0]), int(halo[1])
        keep = []
        if ep is not None:
            sc, sh, relu = ep
            o.ep_scale = sc.data_ptr() if sc is not None else None
            o.ep_shift = sh.data_ptr() if sh is not None else None
            o.relu = 1 if relu else 0
        if y16 is not None:
            buf, off, scl = y16
            o.y16 = buf.data_ptr()
            o.Dy16, o.y16_off, o.y16_scale = int(buf.shape[1]), int(off), float(scl)
        opts = ctypes.byref(o)
    _lib.check(lib.mode_conv3d_ex(_p(x), dtype, _p(w), _p(sample_u), _p(y), n, d, h, wd, k, nout, float(out_scale),
                                  _p(out_scale_dev), _p(bn_sums), lo, hi, impl, opts, _stream()), "mode_conv3d")
    return y


COLL_LOG = None      # diagnostics: set to a list to record the sequence of in-graph collectives (tag, numel)


class ShardSpec:
    """Plane bookkeeping of one D-sharded slab tensor [N, D_local_ext, H, W, C] (see mode_planes_t): owned planes
    [own_lo, own_hi), planes inside the global volume [valid_lo, valid_hi), global voxel count per channel and the
    process group whose ranks hold the other slabs."""

    def __init__(self, own, valid, m_global, group=None):
        self.own, self.valid, self.m_global, self.group = own, valid, int(m_global), group

    def planes(self, rows_per_plane, d):
        return _lib.ModePlanes(rows_per_plane, d, self.own[0], self.own[1], self.valid[0], self.valid[1], self.m_global)

    def world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def all_reduce(self, t, tag=""):
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            if COLL_LOG is not None:
                COLL_LOG.append((tag, t.numel()))
            dist.all_reduce(t, group=self.group)


WGRAD_SPLIT_PHASES = os.environ.get("REPMODE_WGRAD_PHASES", "1") == "1"
K1B_FROM_PARTIALS = os.environ.get("REPMODE_K1B_FROM_PARTIALS", "1") == "1"   # K1b reads K4's partials (no d_weff) when it can


def conv3d_wgrad(x, dy, dtype, n, d, h, wd, ci, co, out_scale_dev=None, impl=0, halo=None, two_phase=False):
    """K4.  d = planes of dy (the OWNED planes); halo = (Dx, x_off) when x carries halo planes.
    two_phase: launch only the tensor-core part now and return (dw, finish): the caller forks K3 -- which needs nothing from
    K4 -- onto the side stream and THEN calls finish() (the deterministic slab reduce that completes dw), so the reduce no
    longer sits between the two tensor-core kernels of the backward pass."""
    lib = _lib.load()
    dw = torch.empty((n, 125, co, ci), dtype=torch.float32, device=x.device)
    impl_eff = impl if impl else (2 if dtype == _lib.MODE_F16 else 1)
    ws_bytes = lib.mode_conv3d_wgrad_workspace_bytes(n, d, h, wd, ci, co, impl_eff)
    ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device)
    dx_, xo = (int(halo[0]), int(halo[1])) if halo is not None else (0, 0)

    def launch(phase):
        _lib.check(lib.mode_conv3d_wgrad_ex(_p(x), _p(dy), dtype, _p(dw), n, d, h, wd, ci, co, 1.0, _p(out_scale_dev),
                                            _p(ws), impl_eff | (phase << 8), dx_, xo, _stream()), "mode_conv3d_wgrad")
    if two_phase and WGRAD_SPLIT_PHASES:
        launch(1)
        # when every unit group ran ONE slab the partials ARE d_weff in another layout: K1b can read them (reparam_bwd)
        lay = (ctypes.c_int32 * 5)()
        rc = lib.mode_conv3d_wgrad_partial_layout(dtype, n, d, h, wd, ci, co, impl_eff, dx_, xo, lay)
        partials = (ws, lay) if (rc == 0 and lay[0] == 1 and lay[1] == 1 and lay[2] == 1) else None
        return dw, (lambda: launch(2)), partials
    launch(0)
    return (dw, (lambda: None), None) if two_phase else dw


def f16_scale_of(tensors, target):
    """Device-side power-of-two scale {s, 1/s} with s*max|t| ~ target over the given fp32 tensors."""
    lib = _lib.load()
    dev = tensors[0].device
    amax = torch.zeros(1, dtype=torch.float32, device=dev)
    if len(tensors) == 1:
        _lib.check(lib.mode_amax(_p(tensors[0]), tensors[0].numel(), _p(amax), _stream()), "mode_amax")
    else:
        k = len(tensors)
        ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in tensors])
        cnts = (ctypes.c_int64 * k)(*[t.numel() for t in tensors])
        _lib.check(lib.mode_amax_multi(ptrs, cnts, k, _p(amax), _stream()), "mode_amax_multi")
    s2 = torch.empty(2, dtype=torch.float32, device=dev)
    _lib.check(lib.mode_f16_scale(_p(amax), float(target), _p(s2), _stream()), "mode_f16_scale")
    return s2


def cast_f16(x, scale_dev=None):
    lib = _lib.load()
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _lib.check(lib.mode_cast_f16(_p(x), _p(out), x.numel(), 1.0, _p(scale_dev), _stream()), "mode_cast_f16")
    return out


def cast_f16_pad(x, c_pad):
    """fp32 [..., C] -> fp16 [..., c_pad] with zero channels appended, one kernel (the stem layer's Ci = 1 -> 32 operand)."""
    lib = _lib.load()
    c = x.shape[-1]
    out = torch.empty(x.shape[:-1] + (c_pad,), dtype=torch.float16, device=x.device)
    _lib.check(lib.mode_cast_f16_pad(_p(x), _p(out), x.numel() // c, c, c_pad, _stream()), "mode_cast_f16_pad")
    return out


_sample_index_cache = {}


def _sample_index(n, dev, identity):
    """arange(n) (train: sample i uses weight slot i) or zeros(n) (eval: everybody uses slot 0) as int32 -- constants, cached
    per (n, device) so that no arange / fill launch sits in front of every layer.  Never cached from inside a CUDA-graph
    capture (the tensor would live in the graph's private pool)."""
    make = lambda: (torch.arange(n, dtype=torch.int32, device=dev) if identity  # noqa: E731
                    else torch.zeros(n, dtype=torch.int32, device=dev))
    if dev.type != "cuda":
        return make()
    key = (n, dev.index if dev.index is not None else torch.cuda.current_device(), identity)
    t = _sample_index_cache.get(key)
    if t is None:
        t = make()
        if not torch.cuda.is_current_stream_capturing():
            _sample_index_cache[key] = t
    return t


def cast_f16_cat(a, b):
    """fp32 [..., Ca] ++ fp32 [..., Cb] -> fp16 [..., Ca + Cb] in one pass (the skip concatenation folded into the cast)."""
    lib = _lib.load()
    ca, cb = a.shape[-1], b.shape[-1]
    out = torch.empty(a.shape[:-1] + (ca + cb,), dtype=torch.float16, device=a.device)
    _lib.check(lib.mode_cast_f16_cat(_p(a), ca, _p(b), cb, _p(out), a.numel() // ca, _stream()), "mode_cast_f16_cat")
    return out


class ModeConvFunction(torch.autograd.Function):
    """MoDEConv.forward as one autograd node.

    forward(x, gate_in, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, running_mean, running_var, training,
            conv_type, precision) -> out [N,Co,D,H,W] (channels_last_3d strides)
    gate_in: int32 task ids [N] (what Net passes) or a float [N,T] embedding (the MoDEConv(x, t) signature).
    """

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    @_on_device_of_first
    def forward(ctx, x, gate_in, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, running_mean, running_var, training,
                conv_type, precision, shard=None, eps=BN_EPS, momentum=BN_MOMENTUM, batches_tracked=None, prebuilt=None,
                x2=None):
        """x2 (optional): a second input concatenated to x along the channels -- MoDEConv(cat((x, x2), 1), t), the decoder's
        skip connection (RepMode.py:106) -- without the concatenated fp32 tensor ever being written."""
        _require_cuda(x, gate_in, k5)
        lib = _lib.load()
        n, ci_x, d, h, wd = x.shape
        layer, ci, co = _layer(k5, k3, k1, a3, a5, gate_w, gate_b)
        c1 = ci_x
        if x2 is not None:
            ci_x += x2.shape[1]
        if ci_x != ci:
            raise RuntimeError(f"MoDEConv: input has {ci_x} channels, layer expects {ci}")
        dev = x.device
        normal = conv_type == "normal"
        if gate_in.dtype.is_floating_point:
            gate_in = gate_in.contiguous().float()
        else:
            gate_in = gate_in.to(torch.int32).contiguous()
        if training:
            U, gate_u = n, gate_in
        else:                                   # eval: the whole batch uses sample 0's kernel (RepMode.py:209-210)
            U, gate_u = 1, gate_in[:1].contiguous()
        sample_u = _sample_index(n, dev, training)
        needs_dx = ctx.needs_input_grad[0] or (x2 is not None and ctx.needs_input_grad[-1])
        needs_dw = any(ctx.needs_input_grad[2:9])
        use_umma = precision == "f16" and umma_shape_ok(ci, co, d, h, wd)
        dtype = _lib.MODE_F16 if use_umma else _lib.MODE_F32

        fold_cat = x2 is not None and use_umma and c1 % 4 == 0 and (ci - c1) % 4 == 0 and ci % 32 == 0
        if x2 is not None and not fold_cat:
            x = torch.cat((x, x2), 1)                   # fp32 SIMT path / odd widths: the plain concatenation
        xn = to_ndhwc(x)
        x2n = to_ndhwc(x2) if fold_cat else None
        w_s2 = None
        ci_p, co_p = (_pad32(ci), _pad32(co)) if use_umma else (ci, co)
        w_scale = W_SCALE_F16 if use_umma else 1.0
        # K1 (re-param, ~10 us of latency-bound work on a small layer) does not depend on x: it runs on the side stream
        # while the main stream stages the fp16 operand
        k1_fork = _Fork(dev, use_umma)
        bn_train = normal and training
        wants_grad = needs_dx or needs_dw or (normal and (ctx.needs_input_grad[9] or ctx.needs_input_grad[10]))
        # the buffers that must be ZERO before their kernels accumulate into them -- the BatchNorm sums of K2's epilogue and
        # the workspace of the BatchNorm-backward reduction -- are one allocation cleared by one fill on the side stream,
        # next to K1; the running-statistics step counter is bumped there too (nothing of this on the critical path)
        sums = bwd_ws = zbuf = None
        if bn_train:
            n_sums = 2 * co_p * 8
            n_ws = int(lib.mode_bn_bwd_workspace_bytes(co)) if (wants_grad and shard is None) else 0
            zbuf = torch.empty(n_sums + n_ws, dtype=torch.uint8, device=dev)
            sums = zbuf[:n_sums].view(torch.float64)
            bwd_ws = zbuf[n_sums:] if n_ws else None

        def side_work():
            if zbuf is not None:
                zbuf.zero_()
            if batches_tracked is not None:
                batches_tracked.add_(1)
        if prebuilt is not None and use_umma and training:
            g, w_fwd, w_dg = prebuilt                    # built by the step's grouped K1 launch (K1Plan); already awaited
            with k1_fork:
                side_work()
        else:
            g, w_fwd, w_dg = reparam_fwd(layer, gate_u, U, ci, co, dtype, needs_dx, w_scale, fork=k1_fork, post=side_work)
        if fold_cat:
            x_op = cast_f16_cat(xn, x2n)
        elif use_umma:
            x_op = cast_f16(xn) if ci_p == ci else cast_f16_pad(xn, ci_p)
        else:
            x_op = xn
        k1_fork.join()

        y = conv3d(x_op, dtype, w_fwd, sample_u, n, d, h, wd, ci_p, co_p, None, sums,
                   stat_range=shard.own if shard is not None else None, out_scale=1.0 / w_scale)
        if co_p != co:                              # drop the zero-padded output channels (head layer, Co = 1)
            y = y[..., :co].contiguous()
            if sums is not None:
                sums = sums.view(2, co_p)[:, :co].contiguous().view(-1)
        m_rows = n * d * h * wd
        planes = shard.planes(h * wd, d) if shard is not None else None
        m_stat = shard.m_global if shard is not None else m_rows
        if shard is not None and bn_train:
            shard.all_reduce(sums, "mode.fwd")            # global statistics over owned voxels of every slab
        mean = invstd = None
        if normal:
            scale = torch.empty(co, dtype=torch.float32, device=dev)
            shift = torch.empty(co, dtype=torch.float32, device=dev)
            out = torch.empty_like(y)
            pl_ref = ctypes.byref(planes) if planes is not None else None
            if training:
                # finalize + apply as ONE kernel (every block derives scale / shift from the sums)
                mean = torch.empty(co, dtype=torch.float32, device=dev)
                invstd = torch.empty(co, dtype=torch.float32, device=dev)
                _lib.check(lib.mode_bn_finalize_apply_relu(_p(sums), m_stat, co, _p(bn_w), _p(bn_b), float(eps),
                                                           float(momentum), _p(mean), _p(invstd), _p(scale), _p(shift),
                                                           _p(running_mean), _p(running_var), _p(y), m_rows, 1, _p(out),
                                                           None, 1.0, pl_ref, None, _stream()), "mode_bn_finalize_apply_relu")
            else:
                invstd_r = torch.rsqrt(running_var + eps)
                scale = (bn_w * invstd_r).contiguous()
                shift = (bn_b - running_mean * scale).contiguous()
                mean, invstd = running_mean.clone(), invstd_r.contiguous()      # frozen statistics (for a backward pass)
                _lib.check(lib.mode_bn_apply_relu(_p(y), m_rows, co, _p(scale), _p(shift), 1, _p(out), None, 1.0, pl_ref,
                                                  _stream()), "mode_bn_apply_relu")
        else:
            out = y
        ctx.frozen_bn = normal and not training
        ctx.shard = shard
        ctx.bwd_ws = bwd_ws if wants_grad else None     # already zero: the backward reduction skips its memset (once)
        if wants_grad:
            x_w = x_op if (UMMA_WGRAD or not use_umma) else xn      # operand K4 will read
            ctx.save_for_backward(None, x_w if needs_dw else None, y if normal else None, g, w_dg,
                                  gate_u, sample_u, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, mean, invstd, w_s2)
            ctx.cfg = (n, d, h, wd, ci, co, U, normal, use_umma, needs_dx, needs_dw, ci_p, co_p)
            ctx.split = c1 if x2 is not None else None
        return from_ndhwc(out)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @_on_device_of_first
    def backward(ctx, dout):
        lib = _lib.load()
        (x_op, x_w, y, g, w_dg, gate_u, sample_u, k5, k3, k1, a3, a5, gate_w, gate_b, bn_w, bn_b, mean, invstd,
         w_s2) = ctx.saved_tensors
        n, d, h, wd, ci, co, U, normal, use_umma, needs_dx, needs_dw, ci_p, co_p = ctx.cfg
        dev = dout.device
        dtype = _lib.MODE_F16 if use_umma else _lib.MODE_F32
        doutn = to_ndhwc(dout)
        m_rows = n * d * h * wd
        dgamma = dbeta = None
        dy_s2 = None
        dy32 = None             # fp32 dy for the SIMT wgrad while UMMA_WGRAD is off
        wgrad_f32 = use_umma and not UMMA_WGRAD and needs_dw
        if normal:
            shard = ctx.shard
            dgamma = torch.empty(co, dtype=torch.float32, device=dev)
            dbeta = torch.empty(co, dtype=torch.float32, device=dev)
            ws, ctx.bwd_ws = ctx.bwd_ws, None           # zeroed during the forward; a second backward zeroes its own
            prezeroed = 1 if ws is not None else 0
            if ws is None:
                ws = torch.empty(int(lib.mode_bn_bwd_workspace_bytes(co)), dtype=torch.uint8, device=dev)
            planes = shard.planes(h * wd, d) if shard is not None else None
            if ctx.frozen_bn:
                # eval mode (fine-tuning with frozen BatchNorm, saliency maps): the statistics are constants, so
                # dy = gamma * invstd_running * dz with no mean terms.  That is exactly what the kernels compute for
                # "halo copy" planes (valid, not owned): declare every plane one.  dgamma / dbeta are still the plain sums.
                planes = _lib.ModePlanes(m_rows, 1, 0, 0, 0, 1, m_rows)
            pl = ctypes.byref(planes) if planes is not None else None
            _lib.check(lib.mode_bn_relu_bwd_reduce_v2(_p(y), _p(doutn), m_rows, co, _p(bn_w), _p(bn_b), _p(mean),
                                                      _p(invstd), pl, _p(ws), prezeroed, None, _stream()),
                       "mode_bn_relu_bwd_reduce")
            if shard is not None:
                shard.all_reduce(ws[:16 * co].view(torch.float64), "mode.bwd")      # {sum dz, sum dz*xhat} over every slab
            dy_f32 = None
            if use_umma:
                dy_op = torch.empty((n, d, h, wd, co), dtype=torch.float16, device=dev)
                dy_s2 = torch.empty(2, dtype=torch.float32, device=dev)
                if wgrad_f32:
                    dy32 = torch.empty((n, d, h, wd, co), dtype=torch.float32, device=dev)
                dy_f32 = dy32
            else:
                dy_op = torch.empty((n, d, h, wd, co), dtype=torch.float32, device=dev)
                dy_f32 = dy_op
            _lib.check(lib.mode_bn_relu_bwd_apply(_p(y), _p(doutn), m_rows, co, _p(bn_w), _p(bn_b), _p(mean), _p(invstd),
                                                  _p(dgamma), _p(dbeta), _p(dy_f32), _p(dy_op) if use_umma else None,
                                                  _p(dy_s2), pl, _p(ws), _stream()), "mode_bn_relu_bwd_apply")
            if shard is not None and shard.world() > 1:
                # dgamma/dbeta came out of the all-reduced sums: hand back this rank's share so that the usual
                # data-parallel gradient sum reproduces them exactly once
                dgamma /= shard.world()
                dbeta /= shard.world()
        elif use_umma:
            dy_s2 = f16_scale_of([doutn], 8192.0)
            dy_op = cast_f16(doutn, dy_s2[0:1])
            dy32 = doutn
        else:
            dy_op = doutn

        if use_umma:
            dy_op = pad_channels(dy_op, co_p)
        # Order: K4 (wgrad, fills every SM) first; then K3 (dgrad) on the side stream NEXT TO the K1b chain on the main
        # stream -- the CTA-pair dgrad leaves SMs idle on a single volume (64 clusters on 148 SMs) and K1b / gate
        # backward are small latency-bound kernels, so they hide completely behind it.
        d_weff = None
        finish_wgrad = None
        k4_partials = None
        if needs_dw:
            if wgrad_f32:
                d_weff = conv3d_wgrad(x_w, dy32, _lib.MODE_F32, n, d, h, wd, ci, co, None)
            else:
                # the tensor-core part of K4 now; its slab reduce after K3 has been forked (K3 needs nothing from K4)
                d_weff, finish_wgrad, k4_partials = conv3d_wgrad(x_w, dy_op, dtype, n, d, h, wd, ci_p, co_p,
                                                                 dy_s2[1:2] if use_umma else None, two_phase=True)
                if not (K1B_FROM_PARTIALS and use_umma and ci_p == ci and co_p == co):
                    k4_partials = None
                if k4_partials is not None:
                    finish_wgrad = None                 # no reduce phase, no d_weff: K1b reads the partials
        dx = None
        dg_fork = _Fork(dev, needs_dx and needs_dw)
        if needs_dx:
            dxn = torch.empty((n, d, h, wd, ci_p), dtype=torch.float32, device=dev)
            with dg_fork:
                conv3d(dy_op, dtype, w_dg, sample_u, n, d, h, wd, co_p, ci_p, dy_s2[1:2] if use_umma else None, None,
                       out_scale=(1.0 / W_SCALE_F16) if use_umma else 1.0, out=dxn)
        if finish_wgrad is not None:
            finish_wgrad()
        grads = [None] * 7
        if needs_dw:
            if not wgrad_f32 and (ci_p != ci or co_p != co):
                d_weff = d_weff[:, :, :co, :ci].contiguous()
            layer, _, _ = _layer(k5, k3, k1, a3, a5, gate_w, gate_b)
            outs = [torch.empty_like(t) for t in (k5, k3, k1, a3, a5, gate_w, gate_b)]
            ws = torch.empty(max(int(lib.mode_reparam_bwd_workspace_bytes(ci, co, n)), 16), dtype=torch.uint8, device=dev)
            ids, dense = (gate_u, None) if not gate_u.dtype.is_floating_point else (None, gate_u)
            if k4_partials is not None:
                part, lay = k4_partials
                _lib.check(lib.mode_reparam_bwd_partial(ctypes.byref(layer), _p(ids), _p(dense), U, _p(sample_u), n, _p(g),
                                                        _p(part), lay, 1.0, _p(dy_s2[1:2]), *[_p(o) for o in outs], _p(ws),
                                                        _stream()), "mode_reparam_bwd_partial")
            else:
                _lib.check(lib.mode_reparam_bwd(ctypes.byref(layer), _p(ids), _p(dense), U, _p(sample_u), n, _p(g),
                                                _p(d_weff), *[_p(o) for o in outs], _p(ws), _stream()), "mode_reparam_bwd")
            grads = outs
        dg_fork.join()
        dx2 = None
        if needs_dx:
            if ci_p != ci:
                dxn = dxn[..., :ci].contiguous()
            dx = from_ndhwc(dxn)
            if ctx.split is not None:                   # two inputs: their gradients are the two channel ranges of dx
                dx, dx2 = dx[:, :ctx.split], dx[:, ctx.split:]
        return (dx, None, *grads, dgamma, dbeta, None, None, None, None, None, None, None, None, None, None, dx2)


EVAL_CACHE = os.environ.get("REPMODE_EVAL_CACHE", "1") == "1"


class EvalWeightCache:
    """Eval-mode W_eff of one MoDEConv for EVERY task, built once per parameter version.

    In eval mode W_eff depends only on (parameters, task) but the reference rebuilds it for all N samples on every call
    and keeps sample 0's (RepMode.py:201-202, :209-210).  Here K1 runs once with U = num_tasks
    gate inputs (task ids 0..T-1); a forward then only points the conv at set `task[0]` through the device-side
    `sample_u` table -- no K1 launch, no host sync on the task id.  The cache is dropped as soon as any of the seven
    parameters changes (tensor version counters / storage pointers), so an optimizer step or load_state_dict is safe."""

    def __init__(self):
        self.key = None
        self.w = None

    def invalidate(self):
        """Forget the cached kernels.  Needed only after an update that the version counters cannot see -- an in-place
        write through `param.data` (EMA / SWA style); MoDEConv calls it on every switch to train mode and on
        load_state_dict, the points at which such updates happen in practice."""
        self.key = None
        self.w = None

    def nbytes(self):
        """Bytes held (0.2 GB fp16 / 0.4 GB fp32 per task over the full-width U-Net's 19 layers)."""
        return 0 if self.w is None else self.w.numel() * self.w.element_size()

    def get(self, params, num_tasks, ci, co, dtype, w_scale):
        key = (dtype, float(w_scale)) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self.key:
            layer, _, _ = _layer(*params)
            ids = torch.arange(num_tasks, dtype=torch.int32, device=params[0].device)
            _, self.w, _ = reparam_fwd(layer, ids, num_tasks, ci, co, dtype, False, w_scale)
            self.key = key
        return self.w


EVAL_F16_ACT = os.environ.get("REPMODE_EVAL_F16_ACT", "1") == "1"    # eval: 'normal' layers hand fp16 activations on
EVAL_GRAPH = os.environ.get("REPMODE_EVAL_GRAPH", "1") == "1"        # eval: repeated Net forwards replay one CUDA graph


_AFFINE_CACHE = {}     # id(BatchNorm weight Parameter) -> (weakref to it, {versions + extra: value})


def bn_eval_affine(bn, extra=None):
    """(scale, shift) of a frozen BatchNorm: y * scale + shift == (y - running_mean) / sqrt(running_var + eps) * w + b.
    Cached per BatchNorm (keyed on the weight Parameter OBJECT -- storage addresses get recycled between modules -- plus
    the version counters of its four tensors): an eval forward of the U-Net would otherwise spend ~130 tiny elementwise
    launches on these 26 vector pairs.  `extra` = (key, fn): also cache fn(scale, shift) (the stride-2 layers fold the pair
    into their GEMM weights)."""
    bn_w, bn_b, rm, rv = bn[:4]
    eps = bn[4] if len(bn) > 4 else BN_EPS
    ent = _AFFINE_CACHE.get(id(bn_w))
    if ent is None or ent[0]() is not bn_w:
        wid = id(bn_w)
        ent = _AFFINE_CACHE[wid] = (weakref.ref(bn_w, lambda _r, wid=wid: _AFFINE_CACHE.pop(wid, None)), {})
    per_bn = ent[1]
    key = (float(eps),) + tuple((t.data_ptr(), t._version) for t in (bn_w, bn_b, rm, rv))
    if extra is not None:
        key = key + (extra[0],)
    hit = per_bn.get(key)
    if hit is None:
        if len(per_bn) > 8:
            per_bn.clear()
        with torch.no_grad():
            scale = (bn_w * torch.rsqrt(rv + eps)).contiguous()
            shift = (bn_b - rm * scale).contiguous()
            hit = (scale, shift) if extra is None else extra[1](scale, shift)
        per_bn[key] = hit
    return hit


def mode_conv_eval(x, task_ids, params, bn, conv_type, precision, cache):
    """MoDEConv.forward in eval mode without autograd (Model.predict, fnet_model.py:149-223): cached W_eff per task,
    conv on set task[0] for the whole batch (RepMode.py:209-210), frozen-statistics BatchNorm + ReLU FOLDED INTO K2's
    epilogue (RepMode.py:209-212: y is written once and never re-read).  On the tensor-core path a 'normal' layer returns
    its activation as an fp16 channels_last_3d tensor -- exactly the operand the next conv would have rounded it to -- so a
    chain of layers is one kernel per layer: no cast, no BatchNorm pass.  'final' layers return fp32."""
    _require_cuda(x, task_ids, params[0])
    n, ci_x, d, h, wd = x.shape
    co, ci = params[0].shape[0], params[0].shape[1]
    if ci_x != ci:
        raise RuntimeError(f"MoDEConv: input has {ci_x} channels, layer expects {ci}")
    use_umma = precision == "f16" and umma_shape_ok(ci, co, d, h, wd)
    dtype = _lib.MODE_F16 if use_umma else _lib.MODE_F32
    ci_p, co_p = (_pad32(ci), _pad32(co)) if use_umma else (ci, co)
    w_scale = W_SCALE_F16 if use_umma else 1.0
    w_all = cache.get(params, params[5].shape[1], ci, co, dtype, w_scale)
    xn = x.permute(0, 2, 3, 4, 1)
    if use_umma:
        if xn.dtype == torch.float16 and xn.is_contiguous():
            x_op = xn                                               # the previous layer's fp16 activation IS the operand
        else:
            xf = xn.contiguous().float()
            x_op = cast_f16(xf) if ci_p == ci else cast_f16_pad(xf, ci_p)
        x_op = pad_channels(x_op, ci_p)
    else:
        x_op = xn.contiguous().float()
    sample_u = task_ids.to(torch.int32).reshape(-1)[:1].expand(n).contiguous()
    normal = conv_type == "normal"
    ep = None
    if normal:
        scale, shift = bn_eval_affine(bn, (("pad", co_p), lambda sc, sh: (pad_channels(sc, co_p), pad_channels(sh, co_p))))
        ep = (scale, shift, True)
    if use_umma and normal and EVAL_F16_ACT:
        y16 = torch.empty((n, d, h, wd, co_p), dtype=torch.float16, device=x.device)
        # K > 32: the CTA-pair kernel accumulates the 32-channel chunks in an fp32 buffer (scratch here)
        scratch = torch.empty((n, d, h, wd, co_p), dtype=torch.float32, device=x.device) if ci_p > 32 else None
        conv3d(x_op, dtype, w_all, sample_u, n, d, h, wd, ci_p, co_p, None, None, out_scale=1.0 / w_scale, ep=ep,
               y16=(y16, 0, 1.0), want_y=False, out=scratch)
        y = y16 if co_p == co else y16[..., :co].contiguous()
    else:
        y = conv3d(x_op, dtype, w_all, sample_u, n, d, h, wd, ci_p, co_p, None, None, out_scale=1.0 / w_scale, ep=ep)
        if co_p != co:
            y = y[..., :co].contiguous()
    return from_ndhwc(y)


def mode_conv(x, gate_in, params, bn, training, conv_type="normal", precision=None, shard=None, prebuilt=None, x2=None):
    """Functional MoDEConv. params: (k5,k3,k1,a3,a5,gate_w,gate_b); bn: (weight,bias,running_mean,running_var) or None;
    shard: ShardSpec when x is one D-slab (with halos) of a larger volume."""
    bn_w, bn_b, rm, rv = bn[:4] if bn is not None else (None, None, None, None)
    eps = bn[4] if bn is not None and len(bn) > 4 else BN_EPS
    momentum = bn[5] if bn is not None and len(bn) > 5 and bn[5] is not None else BN_MOMENTUM
    tracked = bn[6] if bn is not None and len(bn) > 6 else None      # num_batches_tracked: bumped off the critical path
    return ModeConvFunction.apply(x, gate_in, *params, bn_w, bn_b, rm, rv, training, conv_type,
                                  precision or default_precision(), shard, eps, momentum, tracked, prebuilt, x2)


def _row_pitch(t, c):
    """t: logical [N, D, H, W, C] tensor.  Returns the row pitch (floats) if t is a channel range of a dense NDHWC tensor
    (rows of c contiguous floats at a constant pitch >= c, pitch % 4 == 0, 16-byte aligned start), else None."""
    if t.dtype != torch.float32 or t.dim() != 5 or t.shape[-1] != c or t.stride(-1) != 1 or (c & 3):
        return None
    p = t.stride(-2)
    n, d, h, w = t.shape[:4]
    if p < c or (p & 3) or t.data_ptr() & 15:
        return None
    if t.stride(2) != w * p or t.stride(1) != h * w * p or t.stride(0) != d * h * w * p:
        return None
    return p


class BnReluFunction(torch.autograd.Function):
    """BatchNorm3d (+ReLU) on an NDHWC tensor with the path's own kernels (mode_bn_*): the non-MoDE BatchNorms of
    the U-Net (`conv_down.1`, `convt.1`, reference RepMode.py:80-84,97-101) so that no layer leaves NDHWC.

    d2s = (n, d, h, w): y is the [n*d*h*w*8, C] result of the transposed stride-2 conv as a GEMM (rows in (voxel, kd, kh, kw)
    order); the output is the NDHWC volume [n, 2d, 2h, 2w, C] -- the depth-to-space scatter happens in the apply kernel's
    stores, the gather of the incoming gradient in the backward kernels' loads, dy comes back in y's own row order (what
    the GEMM's backward wants): no permute copy in either direction.  An incoming gradient that is a channel range of a
    wider NDHWC tensor (the second half of the decoder's concatenated input) is read in place through its row pitch."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    @_on_device_of_first
    def forward(ctx, y, weight, bias, running_mean, running_var, training, shard=None, eps=BN_EPS, momentum=BN_MOMENTUM,
                d2s=None):
        _require_cuda(y)
        lib = _lib.load()
        yn = y.contiguous()                                    # [N,D,H,W,C] fp32  (d2s: [rows, C])
        c = yn.shape[-1]
        m_rows = yn.numel() // c
        m_stat = shard.m_global if shard is not None else m_rows
        ctx.shard = shard
        ctx.d2s = d2s
        dev = yn.device
        scale = torch.empty(c, dtype=torch.float32, device=dev)
        shift = torch.empty(c, dtype=torch.float32, device=dev)
        mean = invstd = None
        omap = None
        if d2s is not None:
            n_, d_, h_, w_ = d2s
            out = torch.empty((n_, 2 * d_, 2 * h_, 2 * w_, c), dtype=torch.float32, device=dev)
            omap = ctypes.byref(_lib.ModeRowMap(1, d_, h_, w_, c))
        else:
            out = torch.empty_like(yn)
        if training:
            sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
            mean = torch.empty(c, dtype=torch.float32, device=dev)
            invstd = torch.empty(c, dtype=torch.float32, device=dev)
            _lib.check(lib.mode_bn_stats(_p(yn), m_rows, c, _p(sums), _stream()), "mode_bn_stats")
            if shard is not None:
                shard.all_reduce(sums, "bn.fwd")                         # every plane of a stride-2 level is owned: plain sum
            _lib.check(lib.mode_bn_finalize_apply_relu(_p(sums), m_stat, c, _p(weight), _p(bias), float(eps), float(momentum),
                                                       _p(mean), _p(invstd), _p(scale), _p(shift), _p(running_mean),
                                                       _p(running_var), _p(yn), m_rows, 1, _p(out), None, 1.0, None, omap,
                                                       _stream()), "mode_bn_finalize_apply_relu")
        else:
            invstd = torch.rsqrt(running_var + eps).contiguous()
            mean = running_mean.clone()
            scale = (weight * invstd).contiguous()
            shift = (bias - running_mean * scale).contiguous()
            if omap is not None:
                raise NotImplementedError("BnReluFunction: the depth-to-space form is a training-path fusion")
            _lib.check(lib.mode_bn_apply_relu(_p(yn), m_rows, c, _p(scale), _p(shift), 1, _p(out), None, 1.0, None,
                                              _stream()), "mode_bn_apply_relu")
        ctx.training = training
        ctx.save_for_backward(yn, weight, bias, mean, invstd)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    @_on_device_of_first
    def backward(ctx, dout):
        lib = _lib.load()
        yn, weight, bias, mean, invstd = ctx.saved_tensors
        c = yn.shape[-1]
        m_rows = yn.numel() // c
        dev = yn.device
        d2s = ctx.d2s
        # the incoming gradient is read IN PLACE when it is a channel range of a dense NDHWC tensor (no .contiguous() copy)
        pitch = _row_pitch(dout, c) if (d2s is not None or dout.dim() == 5) else None
        if pitch is None:
            doutn = dout.contiguous().float()
            pitch = c
        else:
            doutn = dout
        dmap = None
        if d2s is not None or pitch != c:
            n_, d_, h_, w_ = d2s if d2s is not None else (1, 1, 1, 1)
            dmap = ctypes.byref(_lib.ModeRowMap(1 if d2s is not None else 0, d_, h_, w_, pitch))
        dgamma = torch.empty(c, dtype=torch.float32, device=dev)
        dbeta = torch.empty(c, dtype=torch.float32, device=dev)
        dy = torch.empty_like(yn)
        ws = torch.empty(int(lib.mode_bn_bwd_workspace_bytes(c)), dtype=torch.uint8, device=dev)
        shard = ctx.shard
        planes = shard.planes(m_rows // (yn.shape[0] * yn.shape[1]), yn.shape[1]) if shard is not None else None
        if not ctx.training:              # frozen statistics: every plane a "halo copy" -> no mean terms (see ModeConvFunction)
            planes = _lib.ModePlanes(m_rows, 1, 0, 0, 0, 1, m_rows)
        pl = ctypes.byref(planes) if planes is not None else None
        _lib.check(lib.mode_bn_relu_bwd_reduce_v2(_p(yn), _p(doutn), m_rows, c, _p(weight), _p(bias), _p(mean), _p(invstd),
                                                  pl, _p(ws), 0, dmap, _stream()), "mode_bn_relu_bwd_reduce")
        if shard is not None:
            shard.all_reduce(ws[:16 * c].view(torch.float64), "bn.bwd")
        _lib.check(lib.mode_bn_relu_bwd_apply_v2(_p(yn), _p(doutn), m_rows, c, _p(weight), _p(bias), _p(mean), _p(invstd),
                                                 _p(dgamma), _p(dbeta), _p(dy), None, None, pl, _p(ws), dmap, _stream()),
                   "mode_bn_relu_bwd_apply")
        if shard is not None and shard.world() > 1:
            dgamma /= shard.world()
            dbeta /= shard.world()
        return dy, dgamma, dbeta, None, None, None, None, None, None, None


D2S_FUSED = os.environ.get("REPMODE_D2S_FUSED", "1") == "1"     # A/B: 0 = depth-to-space as a permute copy


class _tf32_matmul:
    """Scope in which cuBLAS may run fp32 GEMMs on the tensor cores with TF32 operands (10-bit mantissa, fp32 accumulate and
    result: the precision of the rest of the tensor-core path).  Without it the stride-2 GEMMs of a training step run as SIMT
    sgemm: 2.1 ms of a 25 ms step at batch 4 for 0.55 % of the FLOPs (r2n profile)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


class _GemmFunction(torch.autograd.Function):
    """a [M, K] @ b [K, N] with both backward GEMMs, each inside the TF32 scope (the flag is process-global and backward runs
    on the autograd thread, so a scope around the forward call alone would not cover them)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        with _tf32_matmul():
            return a @ b

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        with _tf32_matmul():
            da = g @ b.t() if ctx.needs_input_grad[0] else None
            db = a.t() @ g if ctx.needs_input_grad[1] else None
        return da, db


def _gemm(a, b, precision=None):
    """precision 'f16' (the tensor-core path): TF32 tensor-core GEMM; 'f32': plain fp32 (SIMT sgemm), like the SIMT convs."""
    if a.is_cuda and a.dtype == torch.float32 and (precision or default_precision()) == "f16":
        return _GemmFunction.apply(a, b)
    return a @ b


def down_conv_bn_relu(x, conv_w, bn, training, shard=None, precision=None):
    """Conv3d(k=2, s=2, bias=False) + BatchNorm3d + ReLU (reference RepMode.py:80-84) on NDHWC data: the stride-2
    conv is a plain GEMM on the space-to-depth view ([voxels/8, 8*Ci] @ [8*Ci, Co], cuBLAS), BN+ReLU are the path's
    own kernels.  x: [N,C,D,H,W] any strides -> [N,Co,D/2,H/2,W/2] channels_last_3d."""
    xn = x.permute(0, 2, 3, 4, 1)
    n, d, h, w, c = xn.shape
    co = conv_w.shape[0]
    x8 = xn.reshape(n, d // 2, 2, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, 8 * c)
    wm = conv_w.permute(2, 3, 4, 1, 0).reshape(8 * c, co)                     # [(kd,kh,kw,ci), co]
    if not training and not torch.is_grad_enabled():
        # eval: frozen BatchNorm folded into the GEMM (scale into the weight columns, shift as the bias), ReLU in place;
        # the activation keeps its dtype (fp16 on the tensor-core eval path)
        wf, sh = bn_eval_affine((bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps),
                                (("down", conv_w.data_ptr(), conv_w._version, x8.dtype),
                                 lambda sc, shf: ((wm * sc).to(x8.dtype).contiguous(), shf.to(x8.dtype))))
        y = torch.addmm(sh, x8, wf).relu_()
        return y.view(n, d // 2, h // 2, w // 2, co).permute(0, 4, 1, 2, 3)
    y = _gemm(x8, wm.to(x8.dtype), precision).view(n, d // 2, h // 2, w // 2, co)
    out = BnReluFunction.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, shard, bn.eps,
                               bn.momentum if bn.momentum is not None else BN_MOMENTUM)
    return out.permute(0, 4, 1, 2, 3)


def up_conv_bn_relu(x, convt_w, bn, training, shard=None, precision=None):
    """ConvTranspose3d(k=2, s=2, bias=False) + BatchNorm3d + ReLU (reference RepMode.py:97-101) on NDHWC data:
    [voxels, Ci] @ [Ci, 8*Co] (cuBLAS) followed by the depth-to-space scatter, then the path's BN+ReLU kernels."""
    xn = x.permute(0, 2, 3, 4, 1)
    n, d, h, w, c = xn.shape
    co = convt_w.shape[1]
    wm = convt_w.permute(0, 2, 3, 4, 1).reshape(c, 8 * co)                    # [ci, (kd,kh,kw,co)]
    if not training and not torch.is_grad_enabled():
        wf, sh8 = bn_eval_affine((bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps),
                                 (("up", convt_w.data_ptr(), convt_w._version, xn.dtype),
                                  lambda sc, shf: ((wm.view(c, 8, co) * sc).reshape(c, 8 * co).to(xn.dtype).contiguous(),
                                                   shf.repeat(8).to(xn.dtype))))
        y8 = torch.addmm(sh8, xn.reshape(-1, c), wf).relu_().view(n, d, h, w, 2, 2, 2, co)
        y = y8.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(n, 2 * d, 2 * h, 2 * w, co)
        return y.permute(0, 4, 1, 2, 3)
    y8 = _gemm(xn.reshape(-1, c), wm.to(xn.dtype), precision)                 # [voxels, (kd,kh,kw,co)]
    mom = bn.momentum if bn.momentum is not None else BN_MOMENTUM
    if training and shard is None and y8.is_cuda and co % 4 == 0 and D2S_FUSED:
        # the depth-to-space scatter rides on the BatchNorm apply pass (and its gather on the backward passes)
        out = BnReluFunction.apply(y8.view(-1, co), bn.weight, bn.bias, bn.running_mean, bn.running_var, training, None,
                                   bn.eps, mom, (n, d, h, w))
        return out.permute(0, 4, 1, 2, 3)
    y = y8.view(n, d, h, w, 2, 2, 2, co).permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(n, 2 * d, 2 * h, 2 * w, co)
    out = BnReluFunction.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, shard, bn.eps, mom)
    return out.permute(0, 4, 1, 2, 3)
