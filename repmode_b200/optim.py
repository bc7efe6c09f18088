"""Optimizer step of `Model.do_train_iter` (reference fnet/fnet_model.py:55,111-113) on the B200 path.

`FusedAdam` IS a `torch.optim.Adam` -- same constructor, same `param_groups`, same per-parameter state (`step`, `exp_avg`,
`exp_avg_sq`), so the reference's checkpoints (`optimizer.state_dict()` at fnet_model.py:62, `load_state_dict` at :91)
round-trip unchanged -- whose `step()` is one multi-tensor launch of the path's own kernel (`mode_adam_step`,
csrc/optim.cu) over all 309 parameter tensors instead of ~12 elementwise kernels per tensor.  It also speaks GradScaler's
device-side protocol (`_step_supports_amp_scaling`): `scaler.step(optimizer)` hands over `grad_scale` / `found_inf` as
device tensors and the kernel divides / skips by itself, so the training step has no host synchronisation.

There is no CPU fallback: stepping a CPU parameter raises.
"""
import ctypes
import struct

import torch

from . import lib as _lib


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("repmode_b200.optim.FusedAdam: amsgrad / maximize are not implemented (the reference "
                                      "constructs torch.optim.Adam(params, lr) with defaults, fnet_model.py:55)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False)
        self._step_supports_amp_scaling = True      # GradScaler.step passes grad_scale / found_inf instead of syncing
        self._tables = {}                           # per group: (key, tensors_dev, chunks_dev, n_tensors, n_chunks, keep)

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not torch.is_tensor(st["step"]) or st["step"].device != p.device or st["step"].dtype != torch.float32:
            # a checkpoint written by torch's default (non-fused) Adam keeps `step` as a CPU scalar
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32).to(p.device)
        return st

    def _table(self, gi, params):
        """Device tables of (p, g, m, v, step, n) records and (tensor, offset) chunks, rebuilt only when an address changes
        (gradients re-allocated after zero_grad(set_to_none=True), parameters moved by .to())."""
        recs = []
        for p in params:
            st = self._init_state(p)
            g = p.grad
            if g.is_sparse:
                raise RuntimeError("FusedAdam does not support sparse gradients")
            if g.dtype != torch.float32 or p.dtype != torch.float32:
                raise RuntimeError("FusedAdam: fp32 parameters and gradients only (master weights of the path are fp32)")
            if not (p.is_contiguous() or p.is_contiguous(memory_format=torch.channels_last_3d)) or g.stride() != p.stride() \
                    or st["exp_avg"].stride() != p.stride() or st["exp_avg_sq"].stride() != p.stride():
                # the update is elementwise over the STORAGE: every tensor of a record must be dense with the same layout
                p.grad = g = g.contiguous() if p.is_contiguous() else g.contiguous(memory_format=torch.channels_last_3d)
                if g.stride() != p.stride():
                    raise RuntimeError("FusedAdam: parameter, gradient and state must share one dense layout")
            recs.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         st["step"].data_ptr(), p.numel()))
        key = tuple(recs)
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent
        lib = _lib.load()
        chunk = int(lib.mode_adam_chunk_elems())
        dev = params[0].device
        tb = b"".join(struct.pack("<6q", *r) for r in recs)
        cb = bytearray()
        n_chunks = 0
        for i, r in enumerate(recs):
            for off in range(0, r[5], chunk):
                cb += struct.pack("<iiq", i, 0, off)
                n_chunks += 1
        host_t = torch.frombuffer(bytearray(tb), dtype=torch.uint8).pin_memory()
        host_c = torch.frombuffer(cb, dtype=torch.uint8).pin_memory()
        ent = (key, host_t.to(dev, non_blocking=True), host_c.to(dev, non_blocking=True), len(recs), n_chunks,
               (host_t, host_c))            # the pinned staging buffers stay alive: a captured graph re-reads them
        self._tables[gi] = ent
        return ent

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)        # set by GradScaler.step around this call
        found_inf = getattr(self, "found_inf", None)
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            if dev.type != "cuda" or any(p.device != dev for p in params):
                raise RuntimeError("repmode_b200.optim.FusedAdam: CUDA parameters on one device only (no CPU fallback)")
            _, tdev, cdev, nt, nc, _keep = self._table(gi, params)
            b1, b2 = group["betas"]
            lr = group["lr"]
            if torch.is_tensor(lr):
                lr = float(lr)
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(lib.mode_adam_step(ctypes.c_void_p(tdev.data_ptr()), nt, ctypes.c_void_p(cdev.data_ptr()), nc,
                                              float(lr), float(b1), float(b2), float(group["eps"]),
                                              float(group["weight_decay"]),
                                              ctypes.c_void_p(grad_scale.data_ptr()) if grad_scale is not None else None,
                                              ctypes.c_void_p(found_inf.data_ptr()) if found_inf is not None else None,
                                              stream), "mode_adam_step")
        return loss
