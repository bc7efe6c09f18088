"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): direct comparison with the
imported reference -- Model.predict's sliding-window logic and Gaussian map, and the oracle port on fresh seeds."""
import argparse
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _load(name, rel):
    sys.dont_write_bytecode = True
    os.environ.setdefault("WANDB_MODE", "disabled")
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Dummy(torch.nn.Module):
    def forward(self, x, t):
        return x * 2 + t.view(-1, 1, 1, 1, 1).float()


def test_predict_and_gaussian_match_reference():
    import fnet.fnet_model as ours
    ref = _load("ref_fnet_model", "fnet/fnet_model.py")
    assert np.array_equal(ours.get_gaussian((8, 16, 16)), ref.get_gaussian((8, 16, 16)))
    opts = argparse.Namespace(adopted_datasets=["a", "b"], gpu_ids=-1, batch_size_eval=3)
    a, b = ours.Model(opts, gpu_ids=-1), ref.Model(opts, gpu_ids=-1)
    a.net, b.net = _Dummy(), _Dummy()
    x = torch.randn(1, 1, 21, 40, 37)
    t = torch.tensor([1])
    pa, pb = a.predict(x, t, (8, 16, 16)), b.predict(x, t, (8, 16, 16))
    assert torch.allclose(pa, pb, atol=1e-5)


def test_oracle_port_matches_live_reference_fresh_seed():
    from oracle import mode_torch as otc
    ref = _load("ref_repmode", "fnet/nn_modules/RepMode.py")
    torch.manual_seed(123)
    m = ref.MoDEConv(5, 4, 8, 16).train()
    x = torch.randn(2, 8, 6, 10, 9)
    t = torch.tensor([3, 1])
    onehot = torch.nn.functional.one_hot(t, 4).float()
    y_ref = m(x, onehot)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    p["subsequent_layer.0.running_mean"].zero_()
    p["subsequent_layer.0.running_var"].fill_(1)
    y = otc.mode_conv(p, "", x, t, True)
    assert torch.allclose(y, y_ref, atol=2e-5, rtol=1e-4)
