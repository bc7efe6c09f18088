"""GPU: mode_adam_step (csrc/optim.cu) through repmode_b200.optim.FusedAdam against torch.optim.Adam on the CPU (the optimizer
the reference builds, fnet/fnet_model.py:55), including the GradScaler hand-shake of fnet_model.py:111-113."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(32, 32, 5, 5, 5), (7,), (160, 12), (3, 1, 3, 3, 3), (70001,), (64, 64, 5, 5, 5)]


def _make(device):
    g = torch.Generator().manual_seed(0)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(device)) for s in SHAPES]


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(wd):
    from repmode_b200.optim import FusedAdam
    ref_p, our_p = _make("cpu"), _make("cuda")
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=wd)
    ours = FusedAdam(our_p, lr=1e-3, weight_decay=wd)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        for a, b in zip(ref_p, our_p):
            gr = torch.randn(a.shape, generator=g) * (10.0 ** (step - 2))
            a.grad, b.grad = gr.clone(), gr.cuda()          # fresh gradient tensors every step: the tables follow the addresses
        ref.step()
        ours.step()
    for a, b in zip(ref_p, our_p):
        assert float((a.detach() - b.detach().cpu()).abs().max()) <= 2e-6 * float(a.detach().abs().max()) + 1e-7
        sa, sb = ref.state[a], ours.state[b]
        assert float(sb["step"]) == 4.0
        assert torch.allclose(sa["exp_avg"], sb["exp_avg"].cpu(), rtol=1e-5, atol=1e-9)
        assert torch.allclose(sa["exp_avg_sq"], sb["exp_avg_sq"].cpu(), rtol=1e-5, atol=1e-12)


def test_fused_adam_grad_scaler_protocol():
    """scaler.step(FusedAdam): gradients are divided by the scale on the device; a step with an inf gradient changes nothing
    (parameters, moments, step counters) and halves the scale."""
    from repmode_b200.optim import FusedAdam
    ref_p, our_p = _make("cpu"), _make("cuda")
    ref = torch.optim.Adam(ref_p, lr=1e-3)
    ours = FusedAdam(our_p, lr=1e-3)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    g = torch.Generator().manual_seed(2)
    for a, b in zip(ref_p, our_p):
        gr = torch.randn(a.shape, generator=g)
        a.grad, b.grad = gr.clone(), (gr * 1024.0).cuda()          # what scaler.scale(loss).backward() leaves behind
    ref.step()
    scaler.step(ours)
    scaler.update()
    for a, b in zip(ref_p, our_p):
        assert float((a.detach() - b.detach().cpu()).abs().max()) <= 2e-6 * float(a.detach().abs().max()) + 1e-7
    before = [b.detach().clone() for b in our_p]
    for b in our_p:
        b.grad = torch.full_like(b, 1.0)
    our_p[2].grad[0, 0] = float("inf")
    scaler.step(ours)
    scaler.update()
    assert scaler.get_scale() == 512.0
    for b, b0 in zip(our_p, before):
        assert torch.equal(b.detach(), b0) and float(ours.state[b]["step"]) == 1.0


def test_model_do_train_iter_on_gpu_uses_fused_adam_and_learns():
    """Model.do_train_iter (reference fnet_model.py:96-132) end to end on the GPU: autocast + GradScaler + FusedAdam, the
    reference's return contract, and a loss that goes down on a fixed batch."""
    import argparse
    from fnet.fnet_model import Model
    from repmode_b200.optim import FusedAdam
    torch.manual_seed(0)
    opts = argparse.Namespace(adopted_datasets=["a", "b", "c"], gpu_ids=0, batch_size_eval=1)
    m = Model(opts, nn_module="RepMode", lr=1e-3, gpu_ids=0)
    assert isinstance(m.optimizer, FusedAdam)
    x = torch.randn(2, 1, 16, 32, 32)
    y = torch.randn(2, 1, 16, 32, 32)
    t = torch.tensor([2, 0])
    losses = []
    for _ in range(6):
        out, frame = m.do_train_iter(x, y, t)
        assert out.shape == x.shape and out.device.type == "cpu" and list(frame.columns) == ["dataset", "loss"]
        assert list(frame["dataset"]) == ["c", "a"]
        losses.append(float(frame["loss"].mean()))
    assert all(l == l for l in losses) and losses[-1] < losses[0]
    steps = {float(s["step"]) for s in m.optimizer.state.values()}
    assert len(steps) == 1 and 1.0 <= steps.pop() <= 6.0          # GradScaler may skip early steps while it finds its scale
