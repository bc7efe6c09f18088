"""CPU (-m "not gpu"): host logic of the sliding-window inference path (repmode_b200/predict.py, reference
fnet/fnet_model.py:149-223) -- the window enumeration, and the multi-rank form (gloo, world 2): the windows of ONE volume
dealt out to the ranks, accumulators summed once, every rank returning the complete prediction.  The blend arithmetic of
the product path is CUDA-only (csrc/predict.cu, checked by tests/test_gpu_predict.py); here a torch backend defined in this
file stands in for it."""
import argparse

import torch
import torch.distributed as dist

from repmode_b200 import predict as P
from tests.test_parallel_cpu import _spawn


class TorchBlend:
    """Test stand-in for predict.CudaBlend with the same interface (reference fnet_model.py:207-220 written with slices)."""

    def __init__(self, channels, size, gauss, device):
        self.gauss = gauss.float()
        self.pred_sum = torch.zeros((channels,) + tuple(size))
        self.weight_sum = torch.zeros(tuple(size))

    def add(self, pred, starts):
        for j, (d, h, w) in enumerate(starts):
            pd, ph, pw = pred.shape[-3:]
            g = self.gauss[:pd, :ph, :pw]
            self.pred_sum[:, d:d + pd, h:h + ph, w:w + pw] += pred[j].float() * g
            self.weight_sum[d:d + pd, h:h + ph, w:w + pw] += g

    def accumulators(self):
        return [self.pred_sum, self.weight_sum]

    def result(self):
        return self.pred_sum / self.weight_sum


class _Dummy(torch.nn.Module):
    def forward(self, x, t):
        return x * 2 + t.view(-1, 1, 1, 1, 1).float()


def test_windows_cover_the_volume_and_stay_inside():
    size, patch = (21, 40, 37), (8, 16, 16)
    wins = P.windows(size, patch)
    assert len(wins) == 5 * 4 * 4                          # ceil((len - patch) / (patch / 2) + 1) per axis
    cover = torch.zeros(size)
    for a, b, c in wins:
        assert all(0 <= lo < hi <= n and hi - lo == p for (lo, hi), n, p in zip((a, b, c), size, patch))
        cover[a[0]:a[1], b[0]:b[1], c[0]:c[1]] += 1
    assert cover.min() >= 1
    assert P.windows((4, 16, 16), (8, 16, 16)) == [((0, 4), (0, 16), (0, 16))]      # volume thinner than a patch: clamped


def _predict_sharded(rank, world):
    from fnet.fnet_model import get_gaussian
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 1, 21, 40, 37, generator=g)
    gauss = torch.from_numpy(get_gaussian((8, 16, 16)))
    return P.sliding_window_predict(_Dummy(), x, torch.tensor([1]), (8, 16, 16), 3, gauss, group=dist.group.WORLD,
                                    blend_cls=TorchBlend)


def test_predict_windows_dealt_out_over_two_ranks_match_single_rank():
    from fnet.fnet_model import Model, get_gaussian
    outs = _spawn(_predict_sharded)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 1, 21, 40, 37, generator=g)
    single = P.sliding_window_predict(_Dummy(), x, torch.tensor([1]), (8, 16, 16), 3,
                                      torch.from_numpy(get_gaussian((8, 16, 16))), blend_cls=TorchBlend)
    m = Model(argparse.Namespace(adopted_datasets=["a", "b"], gpu_ids=-1, batch_size_eval=3), gpu_ids=-1)
    m.net = _Dummy()
    mirror = m.predict(x, torch.tensor([1]), (8, 16, 16))                  # the CPU mirror of the reference's loop
    assert torch.allclose(single, mirror, atol=1e-5)
    for o in outs:
        assert o.shape == x.shape and torch.allclose(o, single, atol=1e-5)
    assert torch.equal(outs[0], outs[1])                                   # every rank holds the same complete prediction
