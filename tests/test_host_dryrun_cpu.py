"""CPU (-m "not gpu"): dry run of the HOST side of the path.  The C-ABI library is replaced by a recorder that checks
every call against the ctypes signature table (argument count, pointer-ness) and returns success without touching the
data, so the whole Python glue of MoDEConv.forward / backward -- autograd bookkeeping, saved tensors, buffer shapes,
launch ORDER -- runs here without a GPU.  No arithmetic is checked (that is what the -m gpu parity tests are for); this
exists so that a slip in the host code never costs a GPU-box visit."""
import ctypes

import pytest
import torch

from repmode_b200 import functional as Fm, lib as L
from repmode_b200.nn_modules import MoDEConv

HOST_ONLY = ("_bytes", "_elems", "mode_version", "mode_launch_count", "mode_last_error")


class _Recorder:
    def __init__(self, real):
        self.real = real
        self.calls = []

    def __getattr__(self, name):
        if name not in L.SIGNATURES:
            raise AttributeError(name)
        if name.endswith(HOST_ONLY[:2]) or name in HOST_ONLY[2:]:
            return getattr(self.real, name)
        _, argtypes = L.SIGNATURES[name]

        def call(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments, the ABI takes {len(argtypes)}"
            for a, t in zip(args, argtypes):
                if t is ctypes.c_void_p:
                    assert a is None or isinstance(a, (ctypes.c_void_p, int)) or hasattr(a, "_obj") or \
                        isinstance(a, ctypes.Array), f"{name}: {type(a)} where a pointer is expected"
                elif t in (ctypes.c_int32, ctypes.c_int64, ctypes.c_int):
                    assert isinstance(a, int), f"{name}: {type(a)} where an integer is expected"
                elif t is ctypes.c_float:
                    assert isinstance(a, float), f"{name}: {type(a)} where a float is expected"
            self.calls.append(name)
            return 0
        return call


@pytest.fixture
def dry(monkeypatch):
    rec = _Recorder(L.load())
    monkeypatch.setattr(L, "load", lambda: rec)
    monkeypatch.setattr(Fm, "_require_cuda", lambda *ts: None)
    monkeypatch.setattr(Fm, "_stream", lambda: None)
    monkeypatch.setattr(Fm, "OVERLAP", False)
    return rec


@pytest.mark.parametrize("precision", ["f32", "f16"])
@pytest.mark.parametrize("conv_type,ci,co", [("normal", 32, 32), ("normal", 1, 32), ("final", 32, 1), ("normal", 64, 32)])
def test_train_step_host_glue(dry, precision, conv_type, ci, co):
    m = MoDEConv(5, 4, ci, co, conv_type=conv_type).train()
    m.precision = precision
    x = torch.randn(2, ci, 2, 16, 8, requires_grad=True)
    y = m(x, torch.tensor([1, 3]))
    assert y.shape == (2, co, 2, 16, 8)
    y.backward(torch.randn_like(y))
    assert x.grad is not None and x.grad.shape == x.shape
    for name, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
    c = dry.calls
    assert c.index("mode_reparam_fwd") < c.index("mode_conv3d_ex")
    # backward order: wgrad before dgrad (dgrad overlaps the K1b chain), K1b last
    i_w, i_b = c.index("mode_conv3d_wgrad_ex"), c.index("mode_reparam_bwd")
    i_dgrad = [i for i, n in enumerate(c) if n == "mode_conv3d_ex"][1]
    assert i_w < i_dgrad < i_b


def test_stem_without_input_grad(dry):
    """enc1.conv1: the network input needs no gradient -> no dgrad launch, no dgrad weight pack."""
    m = MoDEConv(5, 4, 1, 32).train()
    y = m(torch.randn(1, 1, 2, 16, 8), torch.tensor([0]))
    y.sum().backward()
    assert dry.calls.count("mode_conv3d_ex") == 1 and "mode_conv3d_wgrad_ex" in dry.calls


@pytest.mark.parametrize("precision", ["f32", "f16"])
def test_eval_cache_host_glue(dry, precision):
    m = MoDEConv(5, 4, 32, 32).eval()
    m.precision = precision
    x = torch.randn(3, 32, 2, 16, 8)
    with torch.no_grad():
        m(x, torch.tensor([2, 0, 1]))
        n_first = dry.calls.count("mode_reparam_fwd")
        m(x, torch.tensor([1, 1, 1]))
        assert dry.calls.count("mode_reparam_fwd") == n_first == 1          # built once for all tasks
        m.expert_conv1x1_conv.add_(1.0)                                     # a parameter update invalidates it
        m(x, torch.tensor([1, 1, 1]))
        assert dry.calls.count("mode_reparam_fwd") == 2
        m(x, torch.nn.functional.one_hot(torch.tensor([1, 1, 1]), 4).float())   # dense embedding: uncached path
        assert dry.calls.count("mode_reparam_fwd") == 3
    m(x, torch.tensor([1, 1, 1]))                                           # grad mode: the autograd path
    assert dry.calls.count("mode_reparam_fwd") == 4


def test_whole_net_host_glue(dry):
    import argparse
    from repmode_b200.nn_modules import Net
    net = Net(argparse.Namespace(adopted_datasets=[0, 1, 2], gpu_ids=-1), mult_chan=32).train()
    x = torch.randn(1, 1, 16, 16, 16)
    y = net(x, torch.tensor([2]))
    assert y.shape == x.shape
    y.sum().backward()
    # K4 is two calls per layer: the tensor-core part, then -- after K3 has been forked -- its slab reduce
    assert dry.calls.count("mode_reparam_fwd") == 19 and dry.calls.count("mode_conv3d_wgrad_ex") == 2 * 19
    assert dry.calls.count("mode_conv3d_ex") == 19 + 18        # no dgrad for the stem
    bwd = [c for c in dry.calls if c in ("mode_conv3d_wgrad_ex", "mode_conv3d_ex")][19:]
    assert bwd[:3] == ["mode_conv3d_wgrad_ex", "mode_conv3d_ex", "mode_conv3d_wgrad_ex"], bwd[:6]   # K4, K3, K4 reduce
    assert all(p.grad is not None for p in net.parameters())


class _FakeStream:
    log = []

    def __init__(self, name="side", device=None):
        self.name = name

    def wait_stream(self, other):
        _FakeStream.log.append((self.name, "waits", other.name))


def test_side_stream_forks_are_joined(dry, monkeypatch):
    """REPMODE_OVERLAP: every fork onto the side stream (K1 next to the operand cast, dgrad next to K1b) is ordered
    after the main stream and joined back into it before its results are used."""
    import contextlib
    monkeypatch.setattr(Fm, "OVERLAP", True)
    monkeypatch.setattr(Fm, "_side_streams", {})
    main = _FakeStream("main")
    _FakeStream.log = []
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: main)
    monkeypatch.setattr(torch.cuda, "Stream", _FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    m = MoDEConv(5, 4, 32, 32).train()
    x = torch.randn(1, 32, 2, 16, 8, requires_grad=True)
    m(x, torch.tensor([1])).sum().backward()
    log = _FakeStream.log
    assert log == [("side", "waits", "main"), ("main", "waits", "side")] * 2, log


@pytest.mark.parametrize("precision", ["f32", "f16"])
def test_sharded_block_host_glue(dry, precision):
    """repmode_b200.sharded.sharded_mode_conv (the D-sharded headline block): haloed operand buffers, conv on the owned
    planes only, exchange points in the order halo(x), BN fwd, BN bwd, halo(dy), gradients."""
    from repmode_b200 import peer, sharded
    from repmode_b200 import lib as L2

    class _Comm(peer.TorchComm):
        def __init__(self):
            super().__init__()
            self.world = 2                    # pretend: record the exchange points without a process group
            self.calls = []

        def all_reduce(self, t, tag=""):
            self.calls.append(("all_reduce", tag, t.numel()))
            return t

        def halo_fill(self, ext, h, tag=""):
            self.calls.append(("halo", tag, tuple(ext.shape), h))
            return ext

    monkeypatch_sharded_lib = L2.load()       # the recorder installed by the fixture
    assert monkeypatch_sharded_lib is dry
    comm = _Comm()
    m = MoDEConv(5, 4, 32, 32).train()
    m.precision = precision
    x = torch.randn(1, 32, 6, 16, 8, requires_grad=True)
    y = sharded.sharded_mode_conv(m, x, torch.tensor([2]), comm, 12, tag="blk")
    assert y.shape == (1, 32, 6, 16, 8)
    y.backward(torch.randn_like(y))
    assert x.grad is not None and all(p.grad is not None and p.grad.shape == p.shape for p in m.parameters())
    kinds = [(c[0], c[1]) for c in comm.calls]
    want = [("halo", "blk.x"), ("all_reduce", "blk.bnf"), ("all_reduce", "blk.bnb")]
    want += [("halo", "blk.dy"), ("all_reduce", "blk.grad")]
    assert kinds == want, kinds
    assert comm.calls[0][2] == (1, 10, 16, 8, 32) and comm.calls[0][3] == 2
    c = dry.calls
    assert c.count("mode_conv3d_ex") == 2 and c.count("mode_conv3d_wgrad_ex") == 2      # K4 + its reduce phase


def test_eval_cache_invalidated_by_train_switch_and_load_state_dict(dry):
    """Updates the tensor version counters cannot see (an in-place write through `param.data`, EMA / SWA style) are covered by
    dropping the cached eval kernels on every switch to train mode and on load_state_dict."""
    m = MoDEConv(5, 4, 32, 32).eval()
    m.precision = "f16"
    x = torch.randn(1, 32, 2, 16, 8)
    t = torch.tensor([1])
    with torch.no_grad():
        m(x, t)
        assert dry.calls.count("mode_reparam_fwd") == 1 and m._eval_cache.key is not None
        m.expert_conv5x5_conv.data.mul_(0.5)            # invisible to _version ...
        m(x, t)
        assert dry.calls.count("mode_reparam_fwd") == 1
        m.train(); m.eval()                             # ... but a train/eval round trip drops the cache
        assert m._eval_cache.key is None and m._eval_cache.nbytes() == 0
        m(x, t)
        assert dry.calls.count("mode_reparam_fwd") == 2
        m.load_state_dict(m.state_dict())
        m(x, t)
        assert dry.calls.count("mode_reparam_fwd") == 3
