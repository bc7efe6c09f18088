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


def test_metric_stats_match_reference():
    """fnet.metric.get_metric_stats: same return shape, keys ('MSE', 'MAE', 'R2' -- the DataFrame columns main.py / eval.py
    read) and values as the reference's sklearn-based implementation (fnet/metric.py:7-34)."""
    import fnet.metric as ours
    ref = _load("ref_metric", "fnet/metric.py")
    g = torch.Generator().manual_seed(5)
    pred, target = torch.randn(1, 6, 9, 7, generator=g), torch.randn(1, 6, 9, 7, generator=g)
    ea, sa = ours.get_metric_stats(pred, target)
    eb, sb = ref.get_metric_stats(pred, target)
    assert isinstance(ea, np.ndarray) and ea.shape == eb.shape and np.allclose(ea, eb)
    assert list(sa) == list(sb) == ["MSE", "MAE", "R2"]
    for k in sb:
        assert abs(sa[k] - float(sb[k])) <= 1e-6 * max(1.0, abs(float(sb[k]))), (k, sa[k], sb[k])


def test_do_eval_iter_columns_match_reference():
    """Model.do_eval_iter returns the DataFrame columns eval.py / main.py index (dataset, path_czi, MSE, MAE, R2)."""
    import fnet.fnet_model as ours
    ref = _load("ref_fnet_model2", "fnet/fnet_model.py")
    opts = argparse.Namespace(adopted_datasets=["a", "b"], gpu_ids=-1, batch_size_eval=2)
    a, b = ours.Model(opts, gpu_ids=-1), ref.Model(opts, gpu_ids=-1)
    a.net, b.net = _Dummy(), _Dummy()
    a.patch_size = b.patch_size = (8, 16, 16)
    x = torch.randn(1, 1, 12, 20, 18)
    tgt = torch.randn(1, 1, 12, 20, 18)
    t = torch.tensor([1])
    info = {"dataset": "b", "path_czi": "x.czi"}
    pa, fa = a.do_eval_iter(x, tgt, t, info)
    pb, fb = b.do_eval_iter(x, tgt, t, info)
    assert torch.allclose(pa, pb, atol=1e-5)
    assert list(fa.columns) == list(fb.columns)
    for c in ("MSE", "MAE", "R2"):
        assert abs(float(fa[c][0]) - float(fb[c][0])) < 1e-5
