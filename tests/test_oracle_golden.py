"""CPU: pins both oracle restatements (numpy closed forms, torch port) to the golden vectors frozen
from the live reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import mode_numpy as onp
from oracle import mode_torch as otc
from tests.util import assert_close, load_golden, params_of

CONV_TRAIN = ["conv_train_small", "conv_train_final", "conv_train_stem", "conv_train_c32", "conv_train_final_c32"]
TOL = 2e-5   # fp32 re-association only (the oracles and the reference all compute in fp32)


@pytest.mark.parametrize("name", CONV_TRAIN + ["conv_eval_small"])
def test_numpy_forward(name):
    d = load_golden(name)
    p = params_of(d)
    co = p["expert_conv5x5_conv"].shape[0]
    g = onp.gate_softmax(p["gate.weight"], p["gate.bias"], d["task"], co)
    assert_close(g, d["g"], 1e-6, "g")
    assert_close(onp.reparam_fwd(p, g), d["w_eff"], 1e-6, "w_eff")
    fwd = onp.mode_conv_forward(p, d["x"], d["task"], bool(d["training"]), str(d["conv_type"]))
    assert_close(fwd["out"].astype(np.float32), d["out"], TOL, "out")


@pytest.mark.parametrize("name", CONV_TRAIN)
def test_numpy_backward(name):
    d = load_golden(name)
    p = params_of(d)
    ct = str(d["conv_type"])
    fwd = onp.mode_conv_forward(p, d["x"], d["task"], True, ct)
    dx, grads, _ = onp.mode_conv_backward(p, d["x"], d["task"], fwd, d["dout"], int(d["num_tasks"]), ct)
    assert_close(dx, d["dx"], 5e-5, "dx")
    for k, v in grads.items():
        assert_close(v, d["grad." + k], 1e-4, "grad " + k)


def _tp(d):
    return {k: torch.from_numpy(v) for k, v in params_of(d).items()}


@pytest.mark.parametrize("name", CONV_TRAIN + ["conv_eval_small", "conv_train_config1"])
def test_torch_port_conv(name):
    d = load_golden(name)
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and "pool" not in k)
         for k, v in _tp(d).items()}
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    t = torch.from_numpy(d["task"])
    training = bool(d["training"])
    sub = int(d["sub"])
    y = otc.mode_conv(p, "", x, t, training, str(d["conv_type"]))
    assert_close(y.detach().numpy()[:, :, ::sub, ::sub, ::sub], d["out"], TOL, "out")
    if training:
        (y * torch.from_numpy(d["dout"])).sum().backward()
        assert_close(x.grad.numpy()[:, :, ::sub, ::sub, ::sub], d["dx"], 5e-5, "dx")
        for k in [k for k in d if k.startswith("grad.")]:
            assert_close(p[k[5:]].grad.numpy(), d[k], 1e-4, k)


@pytest.mark.parametrize("name", ["net_eval_small", "net_train_small"])
def test_torch_port_net(name):
    d = load_golden(name)
    training = bool(d["training"])
    p = {k: v.clone().requires_grad_(training and v.dtype.is_floating_point and "running" not in k and "pool" not in k)
         for k, v in _tp(d).items()}
    x = torch.from_numpy(d["x"])
    t = torch.from_numpy(d["task"])
    y = otc.net_forward(p, x, t, training)
    assert_close(y.detach().numpy(), d["out"], 1e-4, "out")
    if training:
        (y * torch.from_numpy(d["dout"])).sum().backward()
        for k in [k for k in d if k.startswith("grad.")]:
            assert_close(p[k[5:]].grad.numpy(), d[k], 2e-4, k)
