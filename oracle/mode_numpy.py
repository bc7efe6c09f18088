"""ORACLE (test infrastructure -- never imported by the product path).

Pure-numpy restatement of RepMode's MoDE-conv hot path, forward AND closed-form backward, with no
autograd: it is the independent checker for the CUDA kernels' intermediate buffers (W_eff, dW_eff,
expert/gate gradients, conv outputs, BN statistics).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this package.

Reference lines restated (paths relative to /root/reference):
  gate + softmax over experts ............ fnet/nn_modules/RepMode.py:198-200
  centre zero-padding of small kernels .... fnet/nn_modules/RepMode.py:165-169  (trans_kernel)
  expert mixing (routing) ................. fnet/nn_modules/RepMode.py:171-192
  avg-pool constants 1/27, 1/125 in fp32 .. fnet/nn_modules/RepMode.py:161-163
  per-sample conv (train) / w[0] (eval) ... fnet/nn_modules/RepMode.py:204-210
  BatchNorm3d + ReLU ...................... fnet/nn_modules/RepMode.py:146-149, 212

Parity pin: the reference ships no tests or golden vectors, so this restatement is
pinned against the LIVE reference module (imported from /root/reference in the build container) by
tests/golden/make_golden.py; the frozen tensors are in tests/golden/*.npz and
tests/test_oracle_golden.py checks this file against them.
"""
import numpy as np

E = 5          # experts: conv5, conv3, conv1, avg3->conv1, avg5->conv1 (RepMode.py:136-142)
KS = 5         # effective kernel size
BN_EPS = 1e-5  # torch.nn.BatchNorm3d default


def pool_const(k, dtype=np.float32):
    """gen_avgpool_kernel: ones(K,K,K).mul(1.0 / K**3) evaluated in fp32 (RepMode.py:161-163)."""
    return np.float32(1.0 / k ** 3).astype(dtype)


def pad5(kernel):
    """trans_kernel: centre-align a [Co,Ci,K,K,K] kernel inside 5x5x5 zeros (RepMode.py:165-169)."""
    k = kernel.shape[2]
    p = (KS - k) // 2
    return np.pad(kernel, [(0, 0), (0, 0), (p, p), (p, p), (p, p)])


def gate_softmax(gate_w, gate_b, task_ids, co):
    """logit[n, e, o] = gate_w[e*Co + o, task[n]] + gate_b[e*Co + o]; softmax over e (RepMode.py:198-200).

    The reference multiplies a one-hot embedding by the Linear weight; that is column task[n] of gate_w.
    """
    logits = gate_w[:, task_ids].T + gate_b[None, :]            # [N, E*Co]
    logits = logits.reshape(len(task_ids), E, co)
    m = logits.max(axis=1, keepdims=True)
    ex = np.exp(logits - m)
    return ex / ex.sum(axis=1, keepdims=True)


def padded_experts(p):
    """The five experts as 5x5x5 kernels, in gate-slot order 0..4 (RepMode.py:173-181)."""
    dt = p["expert_conv5x5_conv"].dtype
    ones3 = np.full((3, 3, 3), pool_const(3, dt), dtype=dt)
    ones5 = np.full((5, 5, 5), pool_const(5, dt), dtype=dt)
    return [
        p["expert_conv5x5_conv"],
        pad5(p["expert_conv3x3_conv"]),
        pad5(p["expert_conv1x1_conv"]),
        pad5(p["expert_avg3x3_conv"] * ones3[None, None]),
        p["expert_avg5x5_conv"] * ones5[None, None],
    ]


def reparam_fwd(p, g):
    """W_eff[n,o,i,:,:,:] = sum_e g[n,e,o] * pad5(K_e)[o,i]  (RepMode.py:183-190). Returns [N,Co,Ci,5,5,5]."""
    ks = padded_experts(p)
    out = []
    for n in range(g.shape[0]):
        w = ks[0] * g[n, 0][:, None, None, None, None]
        for e in range(1, E):
            w = w + ks[e] * g[n, e][:, None, None, None, None]
        out.append(w)
    return np.stack(out)


def reparam_bwd(p, g, task_ids, d_weff, num_tasks):
    """Closed-form gradient of reparam_fwd (+ gate/softmax) w.r.t. the five experts and the gate Linear.

    dK_e = sum_n g[n,e,o] * crop_e(dW[n]);  dg[n,e,o] = sum_{i,tap} pad5(K_e)[o,i,tap] * dW[n,o,i,tap];
    dlogit = g * (dg - sum_e g*dg);  dgate_b = sum_n dlogit;  dgate_w[:, task[n]] += dlogit[n].
    """
    ks = padded_experts(p)
    n_s, _, co = g.shape
    dt = d_weff.dtype
    grads = {
        "expert_conv5x5_conv": np.zeros_like(p["expert_conv5x5_conv"]),
        "expert_conv3x3_conv": np.zeros_like(p["expert_conv3x3_conv"]),
        "expert_conv1x1_conv": np.zeros_like(p["expert_conv1x1_conv"]),
        "expert_avg3x3_conv": np.zeros_like(p["expert_avg3x3_conv"]),
        "expert_avg5x5_conv": np.zeros_like(p["expert_avg5x5_conv"]),
    }
    dg = np.zeros((n_s, E, co), dtype=dt)
    for n in range(n_s):
        dw = d_weff[n]
        gn = [g[n, e][:, None, None, None, None] for e in range(E)]
        grads["expert_conv5x5_conv"] += gn[0] * dw
        grads["expert_conv3x3_conv"] += gn[1] * dw[:, :, 1:4, 1:4, 1:4]
        grads["expert_conv1x1_conv"] += gn[2] * dw[:, :, 2:3, 2:3, 2:3]
        grads["expert_avg3x3_conv"] += gn[3] * dw[:, :, 1:4, 1:4, 1:4].sum(axis=(2, 3, 4), keepdims=True) * pool_const(3, dt)
        grads["expert_avg5x5_conv"] += gn[4] * dw.sum(axis=(2, 3, 4), keepdims=True) * pool_const(5, dt)
        for e in range(E):
            dg[n, e] = (ks[e] * dw).sum(axis=(1, 2, 3, 4))
    dlogit = g * (dg - (g * dg).sum(axis=1, keepdims=True))         # [N,E,Co]
    dlogit = dlogit.reshape(n_s, E * co)
    grads["gate.bias"] = dlogit.sum(axis=0)
    gw = np.zeros((E * co, num_tasks), dtype=dt)
    for n in range(n_s):
        gw[:, task_ids[n]] += dlogit[n]
    grads["gate.weight"] = gw
    return grads


def conv3d_fwd(x, w):
    """y[o] = sum_{i,tap} w[o,i,tap] * xpad[i, . + tap]  -- 5^3 cross-correlation, stride 1, zero pad 2,
    no bias (F.conv3d(..., padding='same'), RepMode.py:207,210).  x [Ci,D,H,W], w [Co,Ci,5,5,5]."""
    ci, d, h, wd = x.shape
    xp = np.pad(x, [(0, 0), (2, 2), (2, 2), (2, 2)])
    y = np.zeros((w.shape[0], d, h, wd), dtype=np.result_type(x, w))
    for kd in range(KS):
        for kh in range(KS):
            for kw in range(KS):
                y += np.einsum("oc,cdhw->odhw", w[:, :, kd, kh, kw], xp[:, kd:kd + d, kh:kh + h, kw:kw + wd])
    return y


def conv3d_dgrad(dy, w):
    """dx[i] = sum_{o,tap} w[o,i,tap] * dypad[o, . - tap + 4]  (conv with flipped, io-transposed kernel)."""
    wf = np.ascontiguousarray(np.flip(w, axis=(2, 3, 4)).transpose(1, 0, 2, 3, 4))
    return conv3d_fwd(dy, wf)


def conv3d_wgrad(x, dy):
    """dW[o,i,tap] = sum_p dy[o,p] * xpad[i, p + tap]."""
    ci, d, h, wd = x.shape
    co = dy.shape[0]
    xp = np.pad(x, [(0, 0), (2, 2), (2, 2), (2, 2)])
    dw = np.zeros((co, ci, KS, KS, KS), dtype=np.result_type(x, dy))
    for kd in range(KS):
        for kh in range(KS):
            for kw in range(KS):
                dw[:, :, kd, kh, kw] = np.einsum("odhw,cdhw->oc", dy, xp[:, kd:kd + d, kh:kh + h, kw:kw + wd])
    return dw


def bn_relu_train_fwd(y, gamma, beta):
    """BatchNorm3d in training mode (batch statistics over N,D,H,W, biased variance) + ReLU.
    y [N,C,D,H,W]. Returns out, (mean, invstd, xhat)."""
    mean = y.mean(axis=(0, 2, 3, 4))
    var = y.var(axis=(0, 2, 3, 4))
    invstd = 1.0 / np.sqrt(var + BN_EPS)
    xhat = (y - mean[None, :, None, None, None]) * invstd[None, :, None, None, None]
    z = xhat * gamma[None, :, None, None, None] + beta[None, :, None, None, None]
    return np.maximum(z, 0), (mean, var, invstd, xhat)


def bn_relu_train_bwd(dout, out, gamma, invstd, xhat):
    """Backward of ReLU(BN_train(y)): returns dy, dgamma, dbeta."""
    dz = dout * (out > 0)
    m = dz.shape[0] * dz.shape[2] * dz.shape[3] * dz.shape[4]
    dbeta = dz.sum(axis=(0, 2, 3, 4))
    dgamma = (dz * xhat).sum(axis=(0, 2, 3, 4))
    b = lambda v: v[None, :, None, None, None]  # noqa: E731
    dy = b(gamma * invstd) * (dz - b(dbeta) / m - xhat * b(dgamma) / m)
    return dy, dgamma, dbeta


def bn_relu_eval_fwd(y, gamma, beta, running_mean, running_var):
    scale = gamma / np.sqrt(running_var + BN_EPS)
    shift = beta - running_mean * scale
    return np.maximum(y * scale[None, :, None, None, None] + shift[None, :, None, None, None], 0)


def mode_conv_forward(p, x, task_ids, training, conv_type="normal"):
    """MoDEConv.forward (RepMode.py:194-214). p holds the state_dict entries of one MoDEConv as numpy arrays.
    Returns a dict with every intermediate the kernels are checked against."""
    co = p["expert_conv5x5_conv"].shape[0]
    g = gate_softmax(p["gate.weight"], p["gate.bias"], task_ids, co)
    w_eff = reparam_fwd(p, g)
    if training:
        y = np.stack([conv3d_fwd(x[n], w_eff[n]) for n in range(x.shape[0])])
    else:
        y = np.stack([conv3d_fwd(x[n], w_eff[0]) for n in range(x.shape[0])])   # RepMode.py:209-210
    res = {"g": g, "w_eff": w_eff, "y": y}
    if conv_type == "normal":
        if training:
            out, (mean, var, invstd, xhat) = bn_relu_train_fwd(
                y, p["subsequent_layer.0.weight"], p["subsequent_layer.0.bias"])
            res.update(mean=mean, var=var, invstd=invstd, xhat=xhat)
        else:
            out = bn_relu_eval_fwd(y, p["subsequent_layer.0.weight"], p["subsequent_layer.0.bias"],
                                   p["subsequent_layer.0.running_mean"], p["subsequent_layer.0.running_var"])
    else:
        out = y
    res["out"] = out
    return res


def mode_conv_backward(p, x, task_ids, fwd, dout, num_tasks, conv_type="normal"):
    """Closed-form backward of the training-mode MoDEConv.forward. Returns dx and parameter grads."""
    grads = {}
    if conv_type == "normal":
        dy, dgamma, dbeta = bn_relu_train_bwd(dout, fwd["out"], p["subsequent_layer.0.weight"], fwd["invstd"],
                                              fwd["xhat"])
        grads["subsequent_layer.0.weight"] = dgamma
        grads["subsequent_layer.0.bias"] = dbeta
    else:
        dy = dout
    n_s = x.shape[0]
    dx = np.stack([conv3d_dgrad(dy[n], fwd["w_eff"][n]) for n in range(n_s)])
    d_weff = np.stack([conv3d_wgrad(x[n], dy[n]) for n in range(n_s)])
    grads.update(reparam_bwd(p, fwd["g"], task_ids, d_weff, num_tasks))
    return dx, grads, {"dy": dy, "d_weff": d_weff}
