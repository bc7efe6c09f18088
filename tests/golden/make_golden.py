"""Generate the golden fixtures in this directory from the LIVE reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 WANDB_MODE=disabled python tests/golden/make_golden.py

Imports fnet.nn_modules.RepMode from /root/reference (read-only), builds the reference modules under
fixed seeds, runs forward/backward on CPU in fp32 and freezes parameters, inputs, outputs and
gradients as .npz.  The fixtures are what pins oracle/ (and through it the CUDA kernels) to the
reference's behaviour; the GPU box has no /root/reference, so tests there read only these files.
"""
import argparse
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _ref():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import fnet.nn_modules.RepMode as ref_mod  # noqa: E402
    sys.path.pop(0)
    return ref_mod


def _np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}   # copy: state_dict aliases live buffers


def onehot(t, num_tasks):
    e = torch.zeros(len(t), num_tasks)
    e[torch.arange(len(t)), t] = 1
    return e


def conv_case(ref_mod, name, seed, num_tasks, ci, co, shape, tasks, training, conv_type, randomize_bn=True,
              np_inputs=False, sub=1):
    """np_inputs: x / dout come from numpy RandomState(seed) / RandomState(seed+1) (a stream numpy keeps
    stable forever) and are NOT stored; sub: store out / dx subsampled by this stride in D, H, W."""
    torch.manual_seed(seed)
    m = ref_mod.MoDEConv(5, num_tasks, ci, co, kernel_size=5, padding="same", conv_type=conv_type)
    if conv_type == "normal" and randomize_bn:
        bn = m.subsequent_layer[0]
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.5, 0.5)
            bn.running_mean.uniform_(-0.2, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    sd0 = _np(m.state_dict())
    n = len(tasks)
    if np_inputs:
        x = torch.from_numpy(np.random.RandomState(seed).standard_normal((n, ci, *shape)).astype(np.float32))
        x.requires_grad_(True)
    else:
        x = torch.randn(n, ci, *shape, requires_grad=True)
    t = torch.tensor(tasks, dtype=torch.int64)
    m.train(training)
    out = {"task": t.numpy(), "training": np.array(training), "conv_type": np.array(conv_type),
           "num_tasks": np.array(num_tasks), "sub": np.array(sub), "np_inputs": np.array(np_inputs),
           "seed": np.array(seed), "x_shape": np.array(x.shape)}
    if not np_inputs:
        out["x"] = x.detach().numpy().copy()
    # intermediates through the reference's own methods
    g = m.softmax(m.gate(onehot(t, num_tasks)).view((n, 5, co)))
    out["g"] = g.detach().numpy()
    out["w_eff"] = m.routing(g, n).detach().numpy()
    y = m(x, onehot(t, num_tasks))
    out["out"] = y.detach().numpy()[:, :, ::sub, ::sub, ::sub]
    if training:
        if np_inputs:
            r = torch.from_numpy(np.random.RandomState(seed + 1).standard_normal(tuple(y.shape)).astype(np.float32))
        else:
            r = torch.randn_like(y)
            out["dout"] = r.numpy()
        (y * r).sum().backward()
        out["dx"] = x.grad.numpy()[:, :, ::sub, ::sub, ::sub]
        for k, v in m.named_parameters():
            out["grad." + k] = v.grad.numpy()
        out.update({"after." + k: v for k, v in _np(m.state_dict()).items() if "running" in k or "num_batches" in k})
    out.update({"p." + k: v for k, v in sd0.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0 and not k.startswith("p.")})


def net_case(ref_mod, name, seed, mult_chan, tasks_all, shape, tasks, training):
    torch.manual_seed(seed)
    opts = argparse.Namespace(adopted_datasets=tasks_all, gpu_ids=-1)
    net = ref_mod.Net(opts, mult_chan=mult_chan)
    # populate BN running stats with two train-mode passes so eval mode is non-trivial
    net.train()
    with torch.no_grad():
        for _ in range(2):
            net(torch.randn(2, 1, *shape), torch.tensor([0, len(tasks_all) - 1]))
    sd0 = _np(net.state_dict())
    x = torch.randn(len(tasks), 1, *shape)
    t = torch.tensor(tasks, dtype=torch.int64)
    net.train(training)
    out = {"x": x.numpy().copy(), "task": t.numpy(), "training": np.array(training),
           "mult_chan": np.array(mult_chan), "num_tasks": np.array(len(tasks_all))}
    if training:
        y = net(x, t)
        r = torch.randn_like(y)
        (y * r).sum().backward()
        out["dout"] = r.numpy()
        for k, v in net.named_parameters():
            out["grad." + k] = v.grad.numpy()
    else:
        with torch.no_grad():
            y = net(x, t)
    out["out"] = y.detach().numpy()
    out.update({"p." + k: v for k, v in sd0.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "out", out["out"].shape, "params", sum(v.size for k, v in out.items() if k.startswith("p.")))


def main_round2(ref_mod):
    """Cases added in round 2 (generated on their own so that the round-1 fixtures stay byte-identical)."""
    # (h) 32 -> 32 'final' layer (no BatchNorm / ReLU), train mode: isolates the tensor-core dgrad / wgrad / K1b on
    #     real-valued data -- no ReLU mask between the operand rounding and the gradients, so the fp16-operand path is held to
    #     1e-3 DIRECTLY against these fp32 vectors
    conv_case(ref_mod, "conv_train_final_c32", 9, 12, 32, 32, (4, 16, 16), [3, 7], True, "final")


def main():
    torch.set_num_threads(8)
    ref_mod = _ref()
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        main_round2(ref_mod)
        with open(os.path.join(HERE, "PROVENANCE.txt"), "a") as f:
            f.write(f"round 2 additions (make_golden.py round2): conv_train_final_c32; torch {torch.__version__}, "
                    f"threads {torch.get_num_threads()}, numpy {np.__version__}, CPU fp32\n")
        return
    # (a) tiny train-mode block, distinct tasks per sample, ragged spatial dims
    conv_case(ref_mod, "conv_train_small", 1, 3, 4, 8, (6, 7, 9), [2, 0], True, "normal")
    # (b) 'final' head (Co=1, no BN), train mode
    conv_case(ref_mod, "conv_train_final", 2, 4, 8, 1, (5, 6, 8), [1, 3, 3], True, "final")
    # (c) eval mode: the whole batch uses sample 0's kernel even when tasks differ (RepMode.py:209-210)
    conv_case(ref_mod, "conv_eval_small", 3, 3, 8, 8, (6, 8, 8), [1, 2, 0], False, "normal")
    # (d) Ci=1 stem layer
    conv_case(ref_mod, "conv_train_stem", 4, 5, 1, 8, (8, 8, 8), [4, 1], True, "normal")
    # (e) tensor-core sized channels (32 -> 32) on a small volume, train mode: the headline layer shape
    conv_case(ref_mod, "conv_train_c32", 5, 12, 32, 32, (4, 16, 16), [3, 7], True, "normal")
    # (f) BASELINE.json configs[0]: single MoDE block, 1x32x32x32 volume, 16 -> 16 channels
    conv_case(ref_mod, "conv_train_config1", 6, 3, 16, 16, (32, 32, 32), [1], True, "normal", np_inputs=True, sub=2)
    # (g) whole U-Net at reduced width (mult_chan=2) so the fixture stays small; 16^3 is the minimum volume
    net_case(ref_mod, "net_eval_small", 7, 2, list(range(3)), (16, 16, 16), [1, 1], False)
    net_case(ref_mod, "net_train_small", 8, 2, list(range(3)), (32, 32, 32), [2, 0], True)
    main_round2(ref_mod)
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by tests/golden/make_golden.py from {REF} (fnet/nn_modules/RepMode.py)\n"
                f"torch {torch.__version__}, threads {torch.get_num_threads()}, numpy {np.__version__}, CPU fp32\n")


if __name__ == "__main__":
    main()
