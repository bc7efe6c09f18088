"""Sliding-window inference of `Model.predict` (reference fnet/fnet_model.py:149-223) on the B200
path: overlapping patches (stride = half a patch, last window clamped to the border), every batch of patches through the
network in eval mode (one CUDA-graph replay per batch after the second one, cached per-task W_eff), Gaussian-weighted
blending by the path's own kernels (csrc/predict.cu) -- and, with a process group, the windows of ONE volume dealt out to
the ranks ("replicas only": patches are independent; SURVEY.md section 8e caveats) with a single sum of the two
accumulators at the end.

The window arithmetic is host logic shared with the CPU mirror in fnet/fnet_model.py; the blend arithmetic lives behind a
small backend object so that the multi-rank host logic can be exercised on CPU/gloo by the tests (tests/ supplies a torch
backend; the product backend below is CUDA-only and raises on CPU tensors)."""
import ctypes
import math

import torch

from . import lib as _lib


def windows(size, patch_size, overlap=0.5):
    """[(d0, d1), (h0, h1), (w0, w1)] of every patch, in the reference's order (fnet_model.py:157-191): stride =
    ceil(patch * (1 - overlap)), steps = ceil((len - patch) / stride + 1), the last window pulled back inside the volume."""
    per_axis = []
    for length, plen in zip(size, patch_size):
        stride = int(math.ceil(plen * (1 - overlap)))
        steps = int(math.ceil((length - plen) / stride + 1))
        axis = []
        for i in range(max(steps, 1)):
            end = min(i * stride + plen, length)
            axis.append((max(end - plen, 0), end))
        per_axis.append(axis)
    return [(a, b, c) for a in per_axis[0] for b in per_axis[1] for c in per_axis[2]]


class CudaBlend:
    """pred_sum / weight_sum accumulators of one volume on the GPU (mode_blend_accumulate / mode_blend_finalize)."""

    def __init__(self, channels, size, gauss, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("repmode_b200.predict: the blend kernels are CUDA-only (no CPU fallback)")
        self.c, self.size, self.device = int(channels), tuple(int(s) for s in size), torch.device(device)
        self.gauss = gauss.to(device=self.device, dtype=torch.float32).contiguous()
        self.pred_sum = torch.zeros((self.c,) + self.size, dtype=torch.float32, device=self.device)
        self.weight_sum = torch.zeros(self.size, dtype=torch.float32, device=self.device)

    def add(self, pred, starts):
        """pred: [P, C, pd, ph, pw] network output for P windows whose origins are `starts` (list of (d, h, w))."""
        lib = _lib.load()
        pred = pred.float().contiguous()
        p, c, pd, ph, pw = pred.shape
        if c != self.c:
            raise RuntimeError(f"predict: the network returned {c} channels, expected {self.c}")
        st = torch.tensor(starts, dtype=torch.int32).to(self.device, non_blocking=True)
        d, h, w = self.size
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            for p0 in range(0, p, 64):
                pn = min(64, p - p0)
                _lib.check(lib.mode_blend_accumulate(ctypes.c_void_p(pred[p0:p0 + pn].data_ptr()),
                                                     ctypes.c_void_p(st[p0:p0 + pn].data_ptr()),
                                                     ctypes.c_void_p(self.gauss.data_ptr()),
                                                     ctypes.c_void_p(self.pred_sum.data_ptr()),
                                                     ctypes.c_void_p(self.weight_sum.data_ptr()), pn, c, pd, ph, pw,
                                                     int(self.gauss.shape[1]), int(self.gauss.shape[2]), d, h, w, stream),
                           "mode_blend_accumulate")

    def accumulators(self):
        return [self.pred_sum, self.weight_sum]

    def result(self):
        lib = _lib.load()
        out = torch.empty_like(self.pred_sum)
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.mode_blend_finalize(ctypes.c_void_p(self.pred_sum.data_ptr()),
                                               ctypes.c_void_p(self.weight_sum.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                               self.c, self.weight_sum.numel(), stream), "mode_blend_finalize")
        return out


def sliding_window_predict(net, signal, task, patch_size, batch_size, gauss, group=None, blend_cls=CudaBlend):
    """signal [1, C, D, H, W] on the device of `net`, task [1] -> blended prediction [1, C, D, H, W] (same device).

    group: a torch.distributed process group whose ranks all hold the SAME volume and the same weights: rank r then runs
    the windows r, r + world, ... and the accumulators are summed over the group once at the end (2 x volume floats through
    NCCL), so every rank returns the complete prediction.  Blending is linear in the patches, so the result equals the
    single-rank one up to fp32 summation order."""
    if signal.shape[0] != 1:
        raise RuntimeError("predict: one volume per call (the reference's accumulators have the signal's shape)")
    size = tuple(signal.shape[-3:])
    wins = windows(size, patch_size)
    world, rank = 1, 0
    if group is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = wins[rank::world]
    blend = blend_cls(signal.shape[1], size, gauss, signal.device)
    bs = max(1, int(batch_size))
    for k in range(0, len(mine), bs):
        chunk = mine[k:k + bs]
        batch = torch.cat([signal[:, :, a[0]:a[1], b[0]:b[1], c[0]:c[1]] for a, b, c in chunk], dim=0)
        with torch.no_grad():
            out = net(batch, task.expand(len(chunk)))
            if isinstance(out, tuple):
                out = out[0]
        blend.add(out, [(a[0], b[0], c[0]) for a, b, c in chunk])
    if world > 1:
        import torch.distributed as dist
        for t in blend.accumulators():
            dist.all_reduce(t, group=group)
    return blend.result().unsqueeze(0)
