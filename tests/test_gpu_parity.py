"""GPU parity tests proper: the CUDA path (through the C ABI) against the golden vectors frozen from the
live reference and against the oracle on seeded inputs.  Tolerances: fp32 SIMT path 1e-4 (re-association
only); tensor-core path (fp16 operands, fp32 accumulate) 1e-3 = the north-star bound."""
import os

import numpy as np
import pytest
import torch

from tests.util import assert_close, load_golden, params_of

pytestmark = pytest.mark.gpu

# torch's OWN kernels serve as references in a few tests below (conv_transpose3d, batch_norm): keep them true fp32, so that a
# comparison measures this library and not cuDNN's / cuBLAS' TF32 defaults.  (test_gpu_net.py sets the same flags at import;
# without them here this module only passed when collected together with it -- seen in the last GPU visit of round 2, where
# test_up_conv_depth_to_space_inside_batchnorm_matches_permute_copy ran alone against a TF32 conv_transpose3d.)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

CONV_CASES = ["conv_train_small", "conv_train_final", "conv_train_stem", "conv_train_c32", "conv_eval_small",
              "conv_train_config1", "conv_train_final_c32"]


def _build_conv(d, precision):
    from repmode_b200.nn_modules import MoDEConv
    p = params_of(d)
    co, ci = p["expert_conv5x5_conv"].shape[:2]
    m = MoDEConv(5, int(d["num_tasks"]), ci, co, conv_type=str(d["conv_type"]))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}, strict=True)
    m.precision = precision
    return m.cuda()


def _oracle_f16(d):
    """Gradients of the tensor-core path are checked against the oracle with the SAME forward operand rounding
    (fp16 operands, fp32 accumulate): 10-bit-mantissa operands flip a few ReLU masks relative to the pure-fp32
    golden run (measured: dx max_rel 5.7e-2 on conv_train_c32 from that alone), which is a property of the
    precision, not of the kernels; the forward OUTPUT is still held to 1e-3 against the fp32 golden vectors."""
    from oracle import mode_torch as otc
    p = {k: torch.from_numpy(v).clone().requires_grad_(v.dtype.kind == "f" and "running" not in k and "pool" not in k)
         for k, v in params_of(d).items()}
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    y = otc.mode_conv(p, "", x, torch.from_numpy(d["task"]), True, str(d["conv_type"]), operand_f16=True)
    (y * torch.from_numpy(d["dout"])).sum().backward()
    return {"dx": x.grad.numpy(), **{"grad." + k: v.grad.numpy() for k, v in p.items() if v.grad is not None}}


@pytest.mark.parametrize("precision,tol", [("f32", 1e-4), ("f16", 1e-3)])
@pytest.mark.parametrize("name", CONV_CASES)
def test_modeconv_vs_golden(name, precision, tol):
    from repmode_b200 import functional as Fm
    d = load_golden(name)
    m = _build_conv(d, precision)
    training = bool(d["training"])
    m.train(training)
    sub = int(d["sub"])
    x = torch.from_numpy(d["x"]).cuda().requires_grad_(training)
    t = torch.from_numpy(d["task"]).cuda().to(torch.int32)
    y = m(x, t)
    assert y.shape == (x.shape[0], m.out_chan) + tuple(x.shape[2:])
    assert_close(y.detach().cpu().numpy()[:, :, ::sub, ::sub, ::sub], d["out"], tol, "out")
    if training:
        (y * torch.from_numpy(d["dout"]).cuda()).sum().backward()
        ref = d
        tensor_core = precision == "f16" and Fm.umma_shape_ok(m.in_chan, m.out_chan, *x.shape[2:])
        if tensor_core and str(d["conv_type"]) == "normal":
            ref = _oracle_f16(d)
            ref["dx"] = ref["dx"][:, :, ::sub, ::sub, ::sub]
        # a 'final' layer (no BatchNorm / ReLU: conv_train_final_c32 runs dgrad, wgrad and K1b on tcgen05) has no ReLU
        # mask that operand rounding could flip, so its gradients are held to the tolerance DIRECTLY against the fp32
        # golden vectors of the live reference
        assert_close(x.grad.cpu().numpy()[:, :, ::sub, ::sub, ::sub], ref["dx"], tol, "dx")
        named = dict(m.named_parameters())
        for k in [k for k in d if k.startswith("grad.")]:
            assert_close(named[k[5:]].grad.cpu().numpy(), ref[k], tol * 2, k)
        if str(d["conv_type"]) == "normal":
            bn = m.subsequent_layer[0]
            assert_close(bn.running_mean.cpu().numpy(), d["after.subsequent_layer.0.running_mean"], 1e-4, "running_mean")
            assert_close(bn.running_var.cpu().numpy(), d["after.subsequent_layer.0.running_var"], 1e-4, "running_var")


def test_modeconv_dense_embedding_matches_ids():
    """MoDEConv.forward(x, t) with the reference's one-hot float embedding == int task ids."""
    d = load_golden("conv_train_small")
    m = _build_conv(d, "f32").eval()
    x = torch.from_numpy(d["x"]).cuda()
    ids = torch.from_numpy(d["task"]).cuda()
    onehot = torch.nn.functional.one_hot(ids, int(d["num_tasks"])).float()
    with torch.no_grad():
        a, b = m(x, ids), m(x, onehot)
    assert torch.equal(a, b)


def test_reparam_weff_bitexact_vs_golden():
    """K1 in fp32 mode reproduces the reference's W_eff to fp32 rounding of the softmax only."""
    import ctypes
    from repmode_b200 import functional as Fm, lib as L
    d = load_golden("conv_train_c32")
    p = {k: torch.from_numpy(v).cuda() for k, v in params_of(d).items()}
    layer, ci, co = Fm._layer(p["expert_conv5x5_conv"], p["expert_conv3x3_conv"], p["expert_conv1x1_conv"],
                              p["expert_avg3x3_conv"], p["expert_avg5x5_conv"], p["gate.weight"], p["gate.bias"])
    ids = torch.from_numpy(d["task"]).cuda().to(torch.int32)
    g, w_fwd, w_dg = Fm.reparam_fwd(layer, ids, len(ids), ci, co, L.MODE_F32, True)
    assert_close(g.cpu().numpy(), d["g"], 1e-6, "g")
    U = len(ids)
    w = w_fwd.view(U, 125, ci // 32, co, 32).permute(0, 3, 2, 4, 1).reshape(U, co, ci, 5, 5, 5)
    assert_close(w.cpu().numpy(), d["w_eff"], 1e-6, "w_eff")
    wd = w_dg.view(U, 125, co // 32, ci, 32).permute(0, 2, 4, 3, 1).reshape(U, co, ci, 125).flip(-1).reshape(U, co, ci, 5, 5, 5)
    assert_close(wd.cpu().numpy(), d["w_eff"], 1e-6, "w_dgrad")


@pytest.mark.parametrize("precision", ["f32", "f16"])
def test_eval_weight_cache_matches_uncached_and_invalidates(precision, monkeypatch):
    """Eval + no_grad (Model.predict, fnet_model.py:149-223): W_eff of every task is built once per parameter version
    (SURVEY.md 8f-3).  The cached path must equal the per-call path bit for bit, use sample 0's task for the whole batch
    (RepMode.py:209-210), launch no K1 kernel on reuse, and rebuild after a parameter update."""
    from repmode_b200 import functional as Fm, lib as L
    d = load_golden("conv_eval_small")
    m = _build_conv(d, precision).eval()
    x = torch.from_numpy(d["x"]).cuda()
    nt = int(d["num_tasks"])
    lib = L.load()
    with torch.no_grad():
        for task in range(nt):
            ids = torch.tensor([task, (task + 1) % nt][: x.shape[0]] * x.shape[0], device="cuda")[: x.shape[0]]
            monkeypatch.setattr(Fm, "EVAL_CACHE", True)
            a = m(x, ids)
            n0 = lib.mode_launch_count()
            a2 = m(x, ids)
            launches_cached = lib.mode_launch_count() - n0
            monkeypatch.setattr(Fm, "EVAL_CACHE", False)
            n0 = lib.mode_launch_count()
            b = m(x, ids)
            launches_plain = lib.mode_launch_count() - n0
            # the cached path folds BatchNorm + ReLU into K2's epilogue (and hands on fp16 on the tensor-core path): the same
            # fp32 value, rounded once more when the activation is fp16
            assert torch.equal(a.float(), b.to(a.dtype).float()) and torch.equal(a, a2)
            assert launches_cached < launches_plain          # no re-parameterisation launch on a cache hit
        monkeypatch.setattr(Fm, "EVAL_CACHE", True)
        before = m(x, ids).clone()
        m.expert_conv5x5_conv.mul_(1.5)                       # in-place update bumps the version counter
        after = m(x, ids)
        monkeypatch.setattr(Fm, "EVAL_CACHE", False)
        assert not torch.equal(before, after)
        assert torch.equal(after.float(), m(x, ids).to(after.dtype).float())


@pytest.mark.parametrize("dtype_name", ["f32", "f16"])
@pytest.mark.parametrize("ci,co,U", [(128, 160, 3), (32, 32, 1), (64, 34, 2)])
def test_reparam_rows_kernel_bit_identical(dtype_name, ci, co, U, monkeypatch):
    """K1's row-block kernel (2 output channels x a whole 32-channel chunk per block, experts read once for all U gate
    inputs; the default when Ci % 32 == 0) writes the same bits as the per-(o, kd slice) kernel: same expressions, same
    association order, same pack -- and the same dgrad pack built from it."""
    from repmode_b200 import functional as Fm, lib as L
    from repmode_b200.nn_modules import MoDEConv
    torch.manual_seed(5)
    m = MoDEConv(5, 12, ci, co).cuda()
    layer, ci, co = Fm._layer(*m._params())
    ids = torch.tensor([3, 7, 3][:U], device="cuda", dtype=torch.int32)
    dtype = L.MODE_F32 if dtype_name == "f32" else L.MODE_F16
    scale = 1.0 if dtype_name == "f32" else Fm.W_SCALE_F16
    monkeypatch.setenv("REPMODE_K1_ROWS", "0")
    g0, w0, d0 = Fm.reparam_fwd(layer, ids, U, ci, co, dtype, True, scale)
    monkeypatch.setenv("REPMODE_K1_ROWS", "1")
    g1, w1, d1 = Fm.reparam_fwd(layer, ids, U, ci, co, dtype, True, scale)
    assert torch.equal(g0, g1) and torch.equal(w0, w1) and torch.equal(d0, d1)


@pytest.mark.parametrize("ci,co,U", [(128, 160, 3), (32, 32, 1), (64, 34, 2), (1, 32, 2), (32, 1, 1)])
def test_pack_dgrad_vector_kernel_bit_identical(ci, co, U, monkeypatch):
    """The 16-byte-access dgrad-pack kernel (pack_dgrad_h16_kernel, default) == the element-wise one, including packs whose
    rows / columns are zero-padded to 32 (stem Ci = 1, head Co = 1)."""
    from repmode_b200 import functional as Fm, lib as L
    from repmode_b200.nn_modules import MoDEConv
    torch.manual_seed(6)
    m = MoDEConv(5, 12, ci, co).cuda()
    layer, ci, co = Fm._layer(*m._params())
    ids = torch.tensor([3, 7, 3][:U], device="cuda", dtype=torch.int32)
    monkeypatch.setenv("REPMODE_PACK_DGRAD_V1", "1")
    _, w0, d0 = Fm.reparam_fwd(layer, ids, U, ci, co, L.MODE_F16, True, Fm.W_SCALE_F16)
    monkeypatch.setenv("REPMODE_PACK_DGRAD_V1", "0")
    _, w1, d1 = Fm.reparam_fwd(layer, ids, U, ci, co, L.MODE_F16, True, Fm.W_SCALE_F16)
    assert torch.equal(w0, w1) and torch.equal(d0, d1) and float(d1.float().abs().max()) > 0


@pytest.mark.parametrize("ci,co,n", [(32, 32, 1), (64, 48, 3), (128, 32, 10)])
def test_reparam_bwd_register_kernel_matches_slab_kernel(ci, co, n, monkeypatch):
    """K1b's register-form kernel (default when Ci % 32 == 0) against the slab kernel on random d_weff: every expert
    gradient and the gate gradients (through gate_bwd, batches beyond its 8-sample chunk included) to fp32 re-association."""
    import ctypes
    from repmode_b200 import functional as Fm, lib as L
    from repmode_b200.nn_modules import MoDEConv
    torch.manual_seed(8)
    m = MoDEConv(5, 12, ci, co).cuda()
    layer, ci, co = Fm._layer(*m._params())
    lib = L.load()
    ids = (torch.arange(n, device="cuda", dtype=torch.int32) * 5) % 12
    sample_u = torch.arange(n, device="cuda", dtype=torch.int32)
    g = torch.softmax(torch.randn(n, 5, co, device="cuda"), dim=1).contiguous()
    d_weff = torch.randn(n, 125, co, ci, device="cuda")

    def run():
        outs = [torch.empty_like(q) for q in m._params()]
        ws = torch.empty(max(int(lib.mode_reparam_bwd_workspace_bytes(ci, co, n)), 16), dtype=torch.uint8, device="cuda")
        L.check(lib.mode_reparam_bwd(ctypes.byref(layer), Fm._p(ids), None, n, Fm._p(sample_u), n, Fm._p(g), Fm._p(d_weff),
                                     *[Fm._p(o) for o in outs], Fm._p(ws), Fm._stream()), "mode_reparam_bwd")
        torch.cuda.synchronize()
        return outs
    monkeypatch.setenv("REPMODE_K1B_SLAB", "1")
    ref = run()
    monkeypatch.delenv("REPMODE_K1B_SLAB")
    got = run()
    # closed form of the expert gradients (RepMode.py:171-192 differentiated): dk5 = sum_n g0[n, o] * dW[n]
    dk5 = torch.einsum("no,ntoi->oit", g[:, 0], d_weff).reshape(co, ci, 5, 5, 5)
    assert_close(got[0].cpu().numpy(), dk5.cpu().numpy(), 1e-5, "dk5 vs closed form")
    for a, b, k in zip(got, ref, ("dk5", "dk3", "dk1", "da3", "da5", "dgate_w", "dgate_b")):
        assert_close(a.cpu().numpy(), b.cpu().numpy(), 2e-5, k)
    # closed form of the gate gradients (softmax + Linear backward on the given g): dg[n,e,o] = <dW[n,:,o,:], K_e[o]>
    import torch.nn.functional as F
    k5, k3, k1, a3, a5 = [q.detach() for q in m._params()[:5]]
    pad = lambda k: F.pad(k, [(5 - k.shape[-1]) // 2] * 6)  # noqa: E731
    ks = torch.stack([k5, pad(k3), pad(k1), pad(a3.expand(-1, -1, 3, 3, 3) / 27), a5.expand(-1, -1, 5, 5, 5) / 125])
    dg = torch.einsum("ntoi,eoit->neo", d_weff.double(), ks.reshape(5, co, ci, 125).double())
    dl = g.double() * (dg - (g.double() * dg).sum(dim=1, keepdim=True))          # [n, e, o]
    dgb = dl.sum(dim=0).reshape(-1)
    dgw = torch.zeros(5 * co, 12, dtype=torch.float64, device="cuda")
    dgw.index_add_(1, ids.long(), dl.permute(1, 2, 0).reshape(5 * co, n))
    assert_close(got[6].cpu().numpy(), dgb.float().cpu().numpy(), 1e-4, "dgate_b vs closed form")
    assert_close(got[5].cpu().numpy(), dgw.float().cpu().numpy(), 1e-4, "dgate_w vs closed form")


@pytest.mark.parametrize("ci,co,n,dense", [(32, 32, 1, False), (64, 48, 20, False), (32, 80, 9, True), (96, 512, 4, False)])
def test_gate_bwd_block_kernel_bit_identical_to_serial(ci, co, n, dense, monkeypatch):
    """The block-form gate backward (32 channels per block, a warp per sample) against the one-thread-per-channel kernel:
    same summation order, so dgate_w / dgate_b must be bit-identical -- task-id and dense gate inputs, Co not a multiple
    of 32, batches beyond the 8-sample chunk."""
    import ctypes
    from repmode_b200 import functional as Fm, lib as L
    from repmode_b200.nn_modules import MoDEConv
    torch.manual_seed(9)
    m = MoDEConv(5, 12, ci, co).cuda()
    layer, ci, co = Fm._layer(*m._params())
    lib = L.load()
    ids = torch.randint(0, 12, (n,), generator=torch.Generator().manual_seed(n)).to("cuda", torch.int32)   # repeats inside a chunk
    emb = torch.randn(n, 12, device="cuda")
    sample_u = torch.arange(n, device="cuda", dtype=torch.int32)
    g = torch.softmax(torch.randn(n, 5, co, device="cuda"), dim=1).contiguous()
    d_weff = torch.randn(n, 125, co, ci, device="cuda")

    def run():
        outs = [torch.full_like(q, float("nan")) for q in m._params()]
        ws = torch.empty(max(int(lib.mode_reparam_bwd_workspace_bytes(ci, co, n)), 16), dtype=torch.uint8, device="cuda")
        L.check(lib.mode_reparam_bwd(ctypes.byref(layer), None if dense else Fm._p(ids), Fm._p(emb) if dense else None, n,
                                     Fm._p(sample_u), n, Fm._p(g), Fm._p(d_weff), *[Fm._p(o) for o in outs], Fm._p(ws),
                                     Fm._stream()), "mode_reparam_bwd")
        torch.cuda.synchronize()
        return outs
    monkeypatch.setenv("REPMODE_GATE_BWD_SERIAL", "1")
    ref = run()
    monkeypatch.delenv("REPMODE_GATE_BWD_SERIAL")
    got = run()
    assert torch.equal(got[5], ref[5]) and torch.equal(got[6], ref[6])
    assert bool(torch.isfinite(got[5]).all()) and bool(torch.isfinite(got[6]).all())


@pytest.mark.parametrize("m_rows,c", [(4096, 32), (1000, 48), (513, 256)])
def test_bn_finalize_apply_one_kernel_matches_two(m_rows, c):
    """mode_bn_finalize_apply_relu (finalize folded into the apply kernel's prologue) against mode_bn_finalize followed by
    mode_bn_apply_relu: output, saved statistics and running statistics bit-identical."""
    from repmode_b200 import functional as Fm, lib as L
    lib = L.load()
    torch.manual_seed(3)
    y = (torch.randn(m_rows, c, device="cuda") * 2 + 0.5).contiguous()
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    L.check(lib.mode_bn_stats(Fm._p(y), m_rows, c, Fm._p(sums), Fm._stream()), "mode_bn_stats")

    def bufs():
        return [torch.empty(c, device="cuda") for _ in range(4)] + [torch.full((c,), 0.25, device="cuda"),
                                                                    torch.full((c,), 1.5, device="cuda")]
    a, b = bufs(), bufs()
    out_a, out_b = torch.empty_like(y), torch.empty_like(y)
    L.check(lib.mode_bn_finalize(Fm._p(sums), m_rows, c, Fm._p(gamma), Fm._p(beta), 1e-5, 0.1, *[Fm._p(t) for t in a],
                                 Fm._stream()), "mode_bn_finalize")
    L.check(lib.mode_bn_apply_relu(Fm._p(y), m_rows, c, Fm._p(a[2]), Fm._p(a[3]), 1, Fm._p(out_a), None, 1.0, None,
                                   Fm._stream()), "mode_bn_apply_relu")
    L.check(lib.mode_bn_finalize_apply_relu(Fm._p(sums), m_rows, c, Fm._p(gamma), Fm._p(beta), 1e-5, 0.1,
                                            *[Fm._p(t) for t in b], Fm._p(y), m_rows, 1, Fm._p(out_b), None, 1.0, None, None,
                                            Fm._stream()), "mode_bn_finalize_apply_relu")
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
    for ta, tb, name in zip(a, b, ("mean", "invstd", "scale", "shift", "running_mean", "running_var")):
        assert torch.equal(ta, tb), name
    ref = torch.relu(torch.nn.functional.batch_norm(y, None, None, gamma, beta, True, 0.0, 1e-5))
    assert_close(out_b.cpu().numpy(), ref.cpu().numpy(), 1e-5, "fused finalize+apply vs torch batch_norm")


@pytest.mark.parametrize("m_rows,c", [(4096, 32), (1000, 64)])
def test_bn_bwd_scale_inside_apply_matches_scale_kernel(m_rows, c, monkeypatch):
    """The fp16 scale of dy derived inside the BatchNorm-backward apply kernel against the separate one-block scale kernel
    (REPMODE_BN_SCALE_KERNEL=1): the published {scale, 1/scale}, the fp16 dy and dgamma / dbeta bit-identical."""
    from repmode_b200 import functional as Fm, lib as L
    lib = L.load()
    torch.manual_seed(4)
    y = torch.randn(m_rows, c, device="cuda").contiguous()
    dout = (torch.randn(m_rows, c, device="cuda") * 1e-3).contiguous()
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    mean, var = y.mean(0), y.var(0, unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)

    def run():
        dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        dy16 = torch.empty(m_rows, c, dtype=torch.float16, device="cuda")
        s2 = torch.zeros(2, device="cuda")
        ws = torch.empty(int(lib.mode_bn_bwd_workspace_bytes(c)), dtype=torch.uint8, device="cuda")
        L.check(lib.mode_bn_relu_bwd_reduce(Fm._p(y), Fm._p(dout), m_rows, c, Fm._p(gamma), Fm._p(beta), Fm._p(mean),
                                            Fm._p(invstd), None, Fm._p(ws), Fm._stream()), "mode_bn_relu_bwd_reduce")
        L.check(lib.mode_bn_relu_bwd_apply(Fm._p(y), Fm._p(dout), m_rows, c, Fm._p(gamma), Fm._p(beta), Fm._p(mean),
                                           Fm._p(invstd), Fm._p(dgamma), Fm._p(dbeta), None, Fm._p(dy16), Fm._p(s2), None,
                                           Fm._p(ws), Fm._stream()), "mode_bn_relu_bwd_apply")
        torch.cuda.synchronize()
        return dgamma, dbeta, dy16, s2
    monkeypatch.setenv("REPMODE_BN_SCALE_KERNEL", "1")
    ref = run()
    monkeypatch.delenv("REPMODE_BN_SCALE_KERNEL")
    got = run()
    for a, b, name in zip(got, ref, ("dgamma", "dbeta", "dy16", "scale2")):
        assert torch.equal(a, b), name
    assert float(got[3][0]) > 1.0 and float(got[3][0] * got[3][1]) == 1.0      # small gradients are scaled UP, by a power of two


def test_reparam_fwd_grouped_bit_identical_to_per_layer():
    """mode_reparam_fwd_grouped (one K1 + one dgrad-pack launch over several layers, the descriptors as a kernel parameter)
    against mode_reparam_fwd per layer: gates, forward packs and dgrad packs bit-identical; a layer without a dgrad buffer
    in the middle of the group (empty tile range) is skipped by the grouped pack kernel."""
    import ctypes
    from repmode_b200 import functional as Fm, lib as L
    from repmode_b200.nn_modules import MoDEConv
    lib = L.load()
    torch.manual_seed(12)
    shapes = [(32, 32), (32, 64), (64, 64), (128, 32), (96, 160)]
    mods = [MoDEConv(5, 12, ci, co).cuda() for ci, co in shapes]
    U = 3
    ids = torch.tensor([4, 11, 4], device="cuda", dtype=torch.int32)
    ref = []
    for i, m in enumerate(mods):
        layer, ci, co = Fm._layer(*m._params())
        ref.append(Fm.reparam_fwd(layer, ids, U, ci, co, L.MODE_F16, i != 2, 256.0))
    arr = (L.ModeReparamItem * len(mods))()
    got = []
    for i, (a, m) in enumerate(zip(arr, mods)):
        layer, ci, co = Fm._layer(*m._params())
        g = torch.full((U, 5, co), float("nan"), device="cuda")
        w = torch.full((U * lib.mode_packed_weight_elems_f16(ci, co),), float("nan"), dtype=torch.float16, device="cuda")
        wd = (torch.full((U * lib.mode_packed_weight_elems_f16(co, ci),), float("nan"), dtype=torch.float16, device="cuda")
              if i != 2 else None)
        a.layer, a.g_out, a.w_fwd, a.w_dgrad = layer, g.data_ptr(), w.data_ptr(), wd.data_ptr() if wd is not None else None
        got.append((g, w, wd))
    L.check(lib.mode_reparam_fwd_grouped(arr, len(mods), Fm._p(ids), None, U, L.MODE_F16, 256.0, Fm._stream()),
            "mode_reparam_fwd_grouped")
    torch.cuda.synchronize()
    L.poll_error("grouped K1")
    for (g0, w0, d0), (g1, w1, d1), sh in zip(ref, got, shapes):
        assert torch.equal(g0, g1), sh
        assert torch.equal(w0.view(torch.int16), w1.view(torch.int16)), sh
        if d0 is not None:
            assert torch.equal(d0.view(torch.int16), d1.view(torch.int16)), sh


@pytest.mark.parametrize("precision", ["f16", "f32"])
def test_two_input_modeconv_matches_concatenated_input(precision):
    """MoDEConv(x, t, x2) -- the decoder's skip concatenation folded into the operand cast (mode_cast_f16_cat) -- against
    MoDEConv(torch.cat((x, x2), 1), t): output and every gradient (both inputs, all parameters) bit-identical."""
    from repmode_b200.nn_modules import MoDEConv
    torch.manual_seed(14)
    m = MoDEConv(5, 12, 64, 32).cuda().train()
    m.precision = precision
    a = torch.randn(2, 32, 6, 16, 16, device="cuda")
    b = torch.randn(2, 32, 6, 16, 16, device="cuda")
    dout = torch.randn(2, 32, 6, 16, 16, device="cuda")
    t = torch.tensor([3, 8], device="cuda")

    def run(two):
        for p in m.parameters():
            p.grad = None
        x1, x2 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = m(x1, t, x2) if two else m(torch.cat((x1, x2), 1), t)
        y.backward(dout)
        torch.cuda.synchronize()
        return [y.detach().clone(), x1.grad.clone(), x2.grad.clone()] + [p.grad.clone() for p in m.parameters()]
    ref, got = run(False), run(True)
    for i, (r, g) in enumerate(zip(ref, got)):
        assert torch.equal(r, g), i


@pytest.mark.parametrize("via_cat", [False, True])
def test_up_conv_depth_to_space_inside_batchnorm_matches_permute_copy(via_cat, monkeypatch):
    """ConvTranspose3d(k=2, s=2) + BatchNorm + ReLU (RepMode.py:97-101) with the depth-to-space scatter / gather folded into
    the BatchNorm kernels' stores / loads (mode_rowmap_t) against the same layer with the permute copy, and (via_cat) with
    the incoming gradient arriving as a channel range of a wider tensor -- read in place through its row pitch -- as the
    decoder's concatenation produces it.  Output, dx, dW, dgamma, dbeta to fp32 re-association (the statistics are summed
    in a different row order)."""
    from repmode_b200 import functional as Fm
    torch.manual_seed(15)
    n, c, co, d, h, w = 2, 64, 32, 3, 6, 8
    convt = torch.nn.ConvTranspose3d(c, co, 2, stride=2, bias=False).cuda()
    bn = torch.nn.BatchNorm3d(co).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    x0 = torch.randn(n, d, h, w, c, device="cuda").permute(0, 4, 1, 2, 3)
    skip = torch.randn(n, 2 * d, 2 * h, 2 * w, co, device="cuda").permute(0, 4, 1, 2, 3)
    dout = torch.randn(n, 2 * d, 2 * h, 2 * w, 2 * co if via_cat else co, device="cuda").permute(0, 4, 1, 2, 3)

    def run(fused):
        monkeypatch.setattr(Fm, "D2S_FUSED", fused)
        for p in list(convt.parameters()) + list(bn.parameters()):
            p.grad = None
        bn.running_mean.zero_(); bn.running_var.fill_(1.0)
        x = x0.clone().requires_grad_(True)
        y = Fm.up_conv_bn_relu(x, convt.weight, bn, True, precision="f32")
        z = torch.cat((skip, y), 1) if via_cat else y
        (z * dout).sum().backward()
        torch.cuda.synchronize()
        return [y.detach().clone(), x.grad.clone(), convt.weight.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone(),
                bn.running_mean.clone(), bn.running_var.clone()]
    ref, got = run(False), run(True)
    for r, g, name in zip(ref, got, ("out", "dx", "dW", "dgamma", "dbeta", "running_mean", "running_var")):
        assert_close(g.cpu().numpy(), r.cpu().numpy(), 2e-5, name)
    # and against torch's own layer
    x = x0.clone().requires_grad_(True)
    yt = torch.relu(torch.nn.functional.batch_norm(torch.nn.functional.conv_transpose3d(x, convt.weight, stride=2), None, None,
                                                   bn.weight, bn.bias, True, 0.0, bn.eps))
    assert_close(got[0].cpu().numpy(), yt.detach().cpu().numpy(), 1e-4, "out vs torch")


def test_wgrad_two_phase_launch_bit_identical(monkeypatch):
    """K4 launched as two calls (tensor-core kernel, then -- after K3 has been forked onto the side stream -- its slab
    reduce; impl bits 8-9 of mode_conv3d_wgrad_ex) against the single call: every gradient bit-identical."""
    from repmode_b200 import functional as Fm
    from repmode_b200.nn_modules import MoDEConv
    torch.manual_seed(16)
    m = MoDEConv(5, 12, 32, 64).cuda().train()
    x0 = torch.randn(2, 32, 8, 32, 32, device="cuda")
    dout = torch.randn(2, 64, 8, 32, 32, device="cuda")
    t = torch.tensor([5, 0], device="cuda")

    def run(phases):
        monkeypatch.setattr(Fm, "WGRAD_SPLIT_PHASES", phases)
        for p in m.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        m(x, t).backward(dout)
        torch.cuda.synchronize()
        return [x.grad.clone()] + [p.grad.clone() for p in m.parameters()]
    ref, got = run(False), run(True)
    for i, (r, g) in enumerate(zip(ref, got)):
        assert torch.equal(r, g), i


@pytest.mark.parametrize("ci,co,n", [(128, 128, 4), (256, 64, 4), (64, 64, 2)])
def test_k1b_from_k4_partials_bit_identical(ci, co, n, monkeypatch):
    """K1b reading d_weff straight out of K4's work-unit partials (mode_reparam_bwd_partial: taken when the wgrad ran one
    slab per unit group, N * Ci/32 * Co/32 >= 50) against the reduce phase + mode_reparam_bwd: every gradient bit-identical.
    (64, 64, 2) has 8 unit groups -> several slabs per unit: the layout query must send it down the ordinary path."""
    import ctypes
    from repmode_b200 import functional as Fm, lib as L
    from repmode_b200.nn_modules import MoDEConv
    lib = L.load()
    lay = (ctypes.c_int32 * 5)()
    rc = lib.mode_conv3d_wgrad_partial_layout(L.MODE_F16, n, 4, 16, 16, ci, co, 2, 0, 0, lay)
    assert rc == 0 and (list(lay)[:3] == [1, 1, 1]) == (n * (ci // 32) * (co // 32) >= 50), list(lay)
    torch.manual_seed(17)
    m = MoDEConv(5, 12, ci, co).cuda().train()
    x0 = torch.randn(n, ci, 4, 16, 16, device="cuda")
    dout = torch.randn(n, co, 4, 16, 16, device="cuda")
    t = (torch.arange(n, device="cuda") * 5) % 12

    def run(flag):
        monkeypatch.setattr(Fm, "K1B_FROM_PARTIALS", flag)
        for p in m.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        n0 = lib.mode_launch_count()
        m(x, t).backward(dout)
        torch.cuda.synchronize()
        L.poll_error("K1b from partials")
        return [x.grad.clone()] + [p.grad.clone() for p in m.parameters()], lib.mode_launch_count() - n0
    (ref, n_ref), (got, n_got) = run(False), run(True)
    for i, (r, g) in enumerate(zip(ref, got)):
        assert torch.equal(r, g), i
    if list(lay)[:3] == [1, 1, 1]:
        assert n_got == n_ref - 1, (n_got, n_ref)          # the reduce launch is gone


def test_cast_f16_pad_matches_cast_then_pad():
    from repmode_b200 import functional as Fm
    x = torch.randn(2, 3, 8, 16, 1, device="cuda") * 3
    x[0, 0, 0, 0, 0] = 1e6                                         # saturates instead of becoming inf
    got = Fm.cast_f16_pad(x, 32)
    want = Fm.pad_channels(Fm.cast_f16(x), 32)
    assert got.shape == want.shape and torch.equal(got, want) and float(got[0, 0, 0, 0, 0]) == 65504.0
    x5 = torch.randn(37, 5, device="cuda")
    assert torch.equal(Fm.cast_f16_pad(x5, 8), Fm.pad_channels(Fm.cast_f16(x5), 8))


def test_frozen_batchnorm_backward_matches_oracle():
    """Backward through a MoDEConv in EVAL mode (frozen BatchNorm statistics: fine-tuning, saliency maps): dx and every
    parameter gradient, BatchNorm affine included, against the oracle's autograd through F.batch_norm(training=False)."""
    from oracle import mode_torch as otc
    d = load_golden("conv_eval_small")
    m = _build_conv(d, "f32").eval()
    x = torch.from_numpy(d["x"]).cuda().requires_grad_(True)
    t = torch.from_numpy(d["task"]).cuda().to(torch.int32)
    g = torch.Generator().manual_seed(3)
    y = m(x, t)
    dout = torch.randn(y.shape, generator=g)
    (y * dout.cuda()).sum().backward()
    p = {k: torch.from_numpy(v).clone().requires_grad_(v.dtype.kind == "f" and "running" not in k and "pool" not in k)
         for k, v in params_of(d).items()}
    xc = torch.from_numpy(d["x"]).requires_grad_(True)
    yc = otc.mode_conv(p, "", xc, torch.from_numpy(d["task"]), False)
    (yc * dout).sum().backward()
    assert_close(y.detach().cpu().numpy(), yc.detach().numpy(), 1e-4, "out")
    assert_close(x.grad.cpu().numpy(), xc.grad.numpy(), 1e-4, "dx")
    for k, q in m.named_parameters():
        assert_close(q.grad.cpu().numpy(), p[k].grad.numpy(), 2e-4, k)


def test_eval_chain_fp16_activations_match_fp32_activations():
    """Eval + no_grad on the tensor-core path: BatchNorm + ReLU folded into K2's epilogue and fp16 activations handed from
    layer to layer (one kernel per MoDEConv) == the same layers with fp32 activations in between, to fp16 rounding of the
    activation (which the next conv's operand staging applies anyway)."""
    from repmode_b200 import functional as Fm
    from repmode_b200.nn_modules import MoDESubNet2Conv
    torch.manual_seed(2)
    stage = MoDESubNet2Conv(5, 12, 32, 64).cuda().eval()
    for c in (stage.conv1, stage.conv2):
        bn = c.subsequent_layer[0]
        with torch.no_grad():
            bn.running_mean.uniform_(-0.2, 0.2); bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    x = torch.randn(2, 32, 6, 32, 16, device="cuda")
    t = torch.tensor([4, 4], device="cuda")
    with torch.no_grad():
        a = stage(x, t)
        assert a.dtype == torch.float16
        old = Fm.EVAL_F16_ACT
        Fm.EVAL_F16_ACT = False
        try:
            b = stage(x, t)
        finally:
            Fm.EVAL_F16_ACT = old
        assert b.dtype == torch.float32
    assert_close(a.float().cpu().numpy(), b.cpu().numpy(), 1e-3, "fp16 vs fp32 activations")
