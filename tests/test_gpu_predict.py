"""GPU: Model.predict on the B200 path (repmode_b200/predict.py + csrc/predict.cu: batched Gaussian blend kernels) against
the CPU mirror of the reference's sliding-window loop (fnet/fnet_model.py:149-223), and the real network through it."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Dummy(torch.nn.Module):
    def forward(self, x, t):
        return x * 2 + t.view(-1, 1, 1, 1, 1).float()


@pytest.mark.parametrize("size,patch,bs", [((21, 40, 37), (8, 16, 16), 3), ((8, 16, 16), (8, 16, 16), 1),
                                           ((4, 30, 16), (8, 16, 16), 70)])
def test_predict_blend_kernels_match_cpu_mirror(size, patch, bs):
    from fnet.fnet_model import Model
    g = torch.Generator().manual_seed(sum(size))
    x = torch.randn(1, 1, *size, generator=g)
    t = torch.tensor([1])
    opts = argparse.Namespace(adopted_datasets=["a", "b"], gpu_ids=-1, batch_size_eval=bs)
    cpu = Model(opts, gpu_ids=-1)
    cpu.net = _Dummy()
    want = cpu.predict(x, t, patch)
    gpu = Model(argparse.Namespace(adopted_datasets=["a", "b"], gpu_ids=0, batch_size_eval=bs), gpu_ids=0)
    gpu.net = _Dummy()
    got = gpu.predict(x, t, patch)
    assert got.device.type == "cpu" and got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_predict_real_net_equals_patchwise_blend():
    """The full-width network through Model.predict (two 32x128x128 windows along W, graph-replayed eval forward) == the
    network called patch by patch and blended on the host."""
    from fnet.fnet_model import Model, get_gaussian
    torch.manual_seed(0)
    opts = argparse.Namespace(adopted_datasets=list(range(12)), gpu_ids=0, batch_size_eval=2)
    m = Model(opts, nn_module="RepMode", gpu_ids=0)
    x = torch.randn(1, 1, 32, 128, 192)
    t = torch.tensor([5])
    got = m.predict(x, t, (32, 128, 128))
    assert got.shape == x.shape and torch.isfinite(got).all()
    gauss = torch.from_numpy(get_gaussian((32, 128, 128)))
    ps, ws = torch.zeros_like(x), torch.zeros_like(x)
    m.net.eval()
    for w0 in (0, 64):
        with torch.no_grad():
            o = m.net(x[..., w0:w0 + 128].cuda(), t.cuda()).float().cpu()
        ps[..., w0:w0 + 128] += o * gauss
        ws[..., w0:w0 + 128] += gauss
    want = ps / ws
    assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())
