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
        for key in ("exp_avg", "exp_avg_sq"):                      # fused multiply-adds vs torch's separate ops: last-bit noise
            ref_t, got_t = sa[key], sb[key].cpu()
            assert float((ref_t - got_t).abs().max()) <= 2e-6 * float(ref_t.abs().max()), key


def test_fused_adam_grad_scaler_protocol():
    """scaler.step(FusedAdam): gradients are divided by the scale on the device; a step with an inf gradient changes nothing
    (parameters, moments, step counters) and halves the scale."""
    from repmode_b200.optim import FusedAdam
    ref_p, our_p = _make("cpu"), _make("cuda")
    ref = torch.optim.Adam(ref_p, lr=1e-3)
    ours = FusedAdam(our_p, lr=1e-3)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    scaler.scale(torch.ones((), device="cuda"))                    # lazily creates the device-side scale (as scale(loss) does)
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
    m = Model(opts, nn_module="RepMode", lr=1e-4, gpu_ids=0)
    assert isinstance(m.optimizer, FusedAdam)
    x = torch.randn(2, 1, 16, 32, 32)
    y = 0.5 * x + 0.1                                   # a target the network can move towards within a few steps
    t = torch.tensor([2, 0])
    losses = []
    for _ in range(10):
        out, frame = m.do_train_iter(x, y, t)
        assert out.shape == x.shape and out.device.type == "cpu" and list(frame.columns) == ["dataset", "loss"]
        assert list(frame["dataset"]) == ["c", "a"]
        losses.append(float(frame["loss"].mean()))
    assert all(l == l for l in losses) and min(losses[-3:]) < losses[0], losses
    steps = {float(s["step"]) for s in m.optimizer.state.values()}
    assert len(steps) == 1 and 1.0 <= steps.pop() <= 10.0          # GradScaler may skip early steps while it finds its scale


def test_out_of_range_task_id_raises_index_error():
    """Reference: IndexError from the host loop of one_hot_task_embedding (RepMode.py:44-49).  Here: K1 clamps the id and
    raises the device flag (code 50); Model.do_train_iter / predict turn it into IndexError."""
    import argparse
    from fnet.fnet_model import Model
    torch.manual_seed(0)
    m = Model(argparse.Namespace(adopted_datasets=["a", "b", "c"], gpu_ids=0, batch_size_eval=1), nn_module="RepMode", gpu_ids=0)
    x = torch.randn(1, 1, 16, 32, 32)
    with pytest.raises(IndexError, match="task id"):
        m.do_train_iter(x, x, torch.tensor([3]))
    m.do_train_iter(x, x, torch.tensor([2]))            # the flag was consumed: a valid step runs clean afterwards
