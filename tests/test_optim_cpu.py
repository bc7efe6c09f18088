"""CPU (-m "not gpu"): repmode_b200.optim.FusedAdam is a torch.optim.Adam as far as the reference's checkpoint code can tell
(fnet/fnet_model.py:55,62,91): same param_groups / state layout, state_dict round trip with a plain torch.optim.Adam in both
directions; and it has no CPU fallback."""
import pytest
import torch


def _params():
    torch.manual_seed(0)
    return [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(5))]


def test_fused_adam_state_dict_is_interchangeable_with_torch_adam():
    from repmode_b200.optim import FusedAdam
    ps = _params()
    ref = torch.optim.Adam(ps, lr=3e-4)
    for p in ps:
        p.grad = torch.ones_like(p)
    ref.step()
    sd = ref.state_dict()
    ours = FusedAdam(_params(), lr=1e-3)
    assert isinstance(ours, torch.optim.Adam) and ours._step_supports_amp_scaling
    ours.load_state_dict(sd)
    assert ours.param_groups[0]["lr"] == 3e-4
    back = ours.state_dict()
    assert back["state"].keys() == sd["state"].keys()
    for k in sd["state"]:
        assert set(back["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert torch.equal(back["state"][k]["exp_avg"], sd["state"][k]["exp_avg"])
        assert float(back["state"][k]["step"]) == 1.0
    again = torch.optim.Adam(_params(), lr=1e-3)
    again.load_state_dict(back)                     # and a plain Adam takes FusedAdam's state


def test_fused_adam_refuses_cpu_parameters_and_unsupported_flags():
    from repmode_b200.optim import FusedAdam
    ps = _params()
    opt = FusedAdam(ps)
    for p in ps:
        p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    with pytest.raises(NotImplementedError):
        FusedAdam(_params(), amsgrad=True)
