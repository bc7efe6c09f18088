"""ORACLE (test infrastructure -- never imported by the product path).

PyTorch-CPU port of RepMode's MoDE-conv path and of the fnet U-Net that hosts it, written functionally
over a state_dict with the reference's key names.  The reference is itself a PyTorch program whose
arithmetic lives in torch (pinned torch==1.12.1, requirements.txt:15; this image has torch 2.11), so
the port uses the same torch CPU operators (F.conv3d -> oneDNN, F.batch_norm, softmax) and is what
bench.py times as the CPU baseline (kind "port") on the GPU box, where /root/reference does not exist.

Reference lines restated (paths relative to /root/reference):
  MoDEConv.forward ............ fnet/nn_modules/RepMode.py:194-214
  routing / trans_kernel ...... fnet/nn_modules/RepMode.py:165-192
  MoDESubNet2Conv ............. fnet/nn_modules/RepMode.py:111-120
  encoder / decoder blocks .... fnet/nn_modules/RepMode.py:74-108
  Net.forward ................. fnet/nn_modules/RepMode.py:51-71

Parity pin: checked against the live reference through tests/golden/*.npz (tests/test_oracle_golden.py)
and, in the build container only, directly against the imported reference (tests/test_reference_live.py).
"""
import torch
import torch.nn.functional as F

E = 5
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _pad5(k):
    p = (5 - k.shape[-1]) // 2
    return F.pad(k, [p, p, p, p, p, p])


def gate_softmax(gate_w, gate_b, task_ids, co):
    logits = gate_w[:, task_ids].t() + gate_b            # one-hot @ W^T == column gather (RepMode.py:44-49,198)
    return torch.softmax(logits.view(-1, E, co), dim=1)  # RepMode.py:199-200


def reparam(p, prefix, g):
    """W_eff [N,Co,Ci,5,5,5] from the five experts and the gates (RepMode.py:171-192)."""
    k5 = p[prefix + "expert_conv5x5_conv"]
    dt = k5.dtype
    c3 = torch.tensor(1.0 / 27, dtype=torch.float32).to(dt)      # fp32-rounded pool constants (RepMode.py:161-163)
    c5 = torch.tensor(1.0 / 125, dtype=torch.float32).to(dt)
    ks = torch.stack([
        k5,
        _pad5(p[prefix + "expert_conv3x3_conv"]),
        _pad5(p[prefix + "expert_conv1x1_conv"]),
        _pad5(p[prefix + "expert_avg3x3_conv"].expand(-1, -1, 3, 3, 3) * c3),
        p[prefix + "expert_avg5x5_conv"].expand(-1, -1, 5, 5, 5) * c5,
    ])                                                           # [E,Co,Ci,5,5,5]
    return torch.einsum("neo,eoidhw->noidhw", g, ks)


def round_f16_scaled(t, amax, target):
    """Straight-through emulation of the tensor-core path's operand staging: multiply by the power of two
    2^floor(log2(target/amax)), round to fp16 (RN), scale back.  Gradients pass through unchanged."""
    a = float(amax)
    s = 2.0 ** torch.floor(torch.log2(torch.tensor(target / a))).item() if a > 0 else 1.0
    r = ((t * s).half().float() / s)
    return t + (r - t).detach()


def mode_conv(p, prefix, x, task_ids, training, conv_type="normal", update_running=False, operand_f16=False):
    """One MoDEConv (RepMode.py:194-214): gate -> softmax -> re-param -> per-sample conv -> BN+ReLU.

    operand_f16=True emulates the B200 tensor-core path's forward numerics (fp16 conv operands with
    power-of-two scaling, fp32 accumulate); the reference itself runs fp32 on CPU (operand_f16=False)."""
    co = p[prefix + "expert_conv5x5_conv"].shape[0]
    g = gate_softmax(p[prefix + "gate.weight"], p[prefix + "gate.bias"], task_ids, co)
    w = reparam(p, prefix, g)
    if operand_f16:
        w = w + (((w * 256.0).half().float() / 256.0) - w).detach()      # fixed 2^8 scale of the fp16 weight pack
        x = x + (x.half().float() - x).detach()
    if training:
        y = torch.cat([F.conv3d(x[i:i + 1], w[i], padding=2) for i in range(x.shape[0])], dim=0)
    else:
        y = F.conv3d(x, w[0], padding=2)                         # eval uses sample 0's kernel (RepMode.py:209-210)
    if conv_type == "normal":
        y = _bn_relu(p, prefix + "subsequent_layer.0.", y, training, update_running)
    return y


def _bn_relu(p, prefix, y, training, update_running):
    rm, rv = p[prefix + "running_mean"], p[prefix + "running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(y, rm, rv, p[prefix + "weight"], p[prefix + "bias"], training, BN_MOMENTUM, BN_EPS)
    return F.relu(y)


def sub2conv(p, prefix, x, t, training, operand_f16=False):
    x = mode_conv(p, prefix + "conv1.", x, t, training, operand_f16=operand_f16)
    return mode_conv(p, prefix + "conv2.", x, t, training, operand_f16=operand_f16)


def _round_tf32(t):
    """Round to a 10-bit mantissa (TF32 operand precision; round to nearest, ties away), straight-through gradient."""
    bits = t.detach().contiguous().view(torch.int32)
    r = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return t + (r - t).detach()


def net_forward(p, x, task_ids, training, operand_f16=False):
    """Net.forward (RepMode.py:51-71) over a reference-keyed state_dict p.  operand_f16: every MoDEConv with the
    tensor-core path's operand rounding (see mode_conv) and the stride-2 convs with TF32 operands (the tensor-core path runs
    them as TF32 GEMMs in training)."""
    rnd = _round_tf32 if operand_f16 else (lambda v: v)
    skips = []
    for k in (1, 2, 3, 4):
        pre = f"encoder_block{k}."
        s = sub2conv(p, pre + "conv_more.", x, task_ids, training, operand_f16)
        skips.append(s)
        x = F.conv3d(rnd(s), rnd(p[pre + "conv_down.0.weight"]), stride=2)
        x = _bn_relu(p, pre + "conv_down.1.", x, training, False)
    x = sub2conv(p, "bottle_block.", x, task_ids, training, operand_f16)
    for k in (4, 3, 2, 1):
        pre = f"decoder_block{k}."
        x = F.conv_transpose3d(rnd(x), rnd(p[pre + "convt.0.weight"]), stride=2)
        x = _bn_relu(p, pre + "convt.1.", x, training, False)
        x = torch.cat((skips[k - 1], x), dim=1)
        x = sub2conv(p, pre + "conv_less.", x, task_ids, training, operand_f16)
    return mode_conv(p, "conv_out.", x, task_ids, training, conv_type="final", operand_f16=operand_f16)


def init_mode_conv_params(num_tasks, ci, co, generator=None, conv_type="normal", prefix=""):
    """Random parameters with the reference's shapes and init ranges (kaiming_uniform_(a=sqrt(5)) ==
    U(+-1/sqrt(fan_in)), RepMode.py:156-159; Linear/BN defaults).  Used for synthetic benchmarks."""
    def u(shape, bound):
        return (torch.rand(shape, generator=generator) * 2 - 1) * bound
    p = {}
    for name, k in (("expert_conv5x5_conv", 5), ("expert_conv3x3_conv", 3), ("expert_conv1x1_conv", 1),
                    ("expert_avg3x3_conv", 1), ("expert_avg5x5_conv", 1)):
        p[prefix + name] = u((co, ci, k, k, k), 1.0 / (ci * k ** 3) ** 0.5)
    p[prefix + "expert_avg3x3_pool"] = torch.ones(3, 3, 3).mul(1.0 / 27)
    p[prefix + "expert_avg5x5_pool"] = torch.ones(5, 5, 5).mul(1.0 / 125)
    p[prefix + "gate.weight"] = u((E * co, num_tasks), 1.0 / num_tasks ** 0.5)
    p[prefix + "gate.bias"] = u((E * co,), 1.0 / num_tasks ** 0.5)
    if conv_type == "normal":
        q = prefix + "subsequent_layer.0."
        p[q + "weight"] = torch.ones(co)
        p[q + "bias"] = torch.zeros(co)
        p[q + "running_mean"] = torch.zeros(co)
        p[q + "running_var"] = torch.ones(co)
        p[q + "num_batches_tracked"] = torch.tensor(0)
    return p
