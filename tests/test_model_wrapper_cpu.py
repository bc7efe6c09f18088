"""CPU (-m "not gpu"): the `fnet.fnet_model.Model` host wrapper (reference API, fnet/fnet_model.py:16-239) around the
B200 network -- construction by plugin name, checkpoint save / load round trip (fnet_model.py:57-94), optimizer-state
device moves.  The network itself refuses CPU tensors (no CPU fallback), so nothing here runs a forward pass."""
import argparse

import pytest
import torch


def _opts(gpu_ids=-1):
    return argparse.Namespace(adopted_datasets=["dna", "lamin", "tom20"], gpu_ids=gpu_ids, batch_size_eval=2)


def test_model_builds_plugin_by_name_and_round_trips_a_checkpoint(tmp_path):
    from fnet.fnet_model import Model
    torch.manual_seed(0)
    a = Model(_opts(), nn_module="RepMode", lr=3e-4, gpu_ids=-1)
    assert type(a.net).__name__ == "Net" and a.net.num_tasks == 3
    assert a.optimizer.param_groups[0]["lr"] == 3e-4
    # give Adam some state without running the network: fake gradients, one step
    for p in a.net.parameters():
        p.grad = torch.full_like(p, 1e-3)
    a.optimizer.step()
    a.count_iter, a.count_epoch = 17, 3
    path = tmp_path / "sub" / "ckpt.p"                       # save_state creates the directory (fnet_model.py:76-78)
    a.save_state(str(path))
    state = torch.load(str(path), weights_only=False)
    assert set(state) == {"nn_module", "opts", "nn_state", "optimizer_state", "count_iter", "count_epoch"}
    assert len(state["nn_state"]) == 309                     # the reference's key count

    torch.manual_seed(1)
    b = Model(_opts(), nn_module=None, gpu_ids=-1)           # eval.py style: empty shell, everything from the file
    assert b.net is None
    b.load_state(str(path), gpu_ids=-1)
    assert b.nn_module == "RepMode" and (b.count_iter, b.count_epoch) == (17, 3)
    sa, sb = a.net.state_dict(), b.net.state_dict()
    assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    oa, ob = a.optimizer.state_dict()["state"], b.optimizer.state_dict()["state"]
    assert oa.keys() == ob.keys()
    assert all(torch.equal(oa[k]["exp_avg"], ob[k]["exp_avg"]) for k in oa)
    assert "Network:\nRepMode\n" in str(b)


def test_set_gpu_recursive_moves_nested_state():
    from fnet.fnet_model import _set_gpu_recursive
    state = {0: {"step": torch.tensor(3.0), "exp_avg": torch.ones(2), "nested": {"v": torch.zeros(1)}, "n": 5}}
    _set_gpu_recursive(state, -1)
    assert state[0]["exp_avg"].device.type == "cpu" and state[0]["nested"]["v"].device.type == "cpu" and state[0]["n"] == 5


def test_multi_gpu_dataparallel_is_refused():
    """The reference's only multi-GPU path is torch.nn.DataParallel (fnet_model.py:40-44, latently broken); this
    framework shards with one process per GPU instead and says so."""
    from fnet.fnet_model import Model
    with pytest.raises(NotImplementedError, match="one process per GPU"):
        Model(_opts([-1, -1]), nn_module="RepMode", gpu_ids=[-1, -1])
