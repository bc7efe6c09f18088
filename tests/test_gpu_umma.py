"""GPU: the tcgen05 implicit-GEMM conv (mode_conv3d impl=2) against the SIMT fp32 kernel and the numpy
oracle, through the C ABI.  Integer-valued fp16-exact operands make the comparison BIT-EXACT (every product
and partial sum is exactly representable in fp32), so any layout / descriptor / pipeline slip shows up as a
hard mismatch rather than a tolerance question."""
import os

import numpy as np
import pytest
import torch

from oracle import mode_numpy as onp
from tests.util import assert_close, pack_weights

pytestmark = pytest.mark.gpu

SHAPES = [  # N, D, H, W, K, Nout
    (1, 1, 16, 8, 32, 32),
    (2, 4, 16, 8, 32, 32),
    (1, 9, 32, 16, 32, 32),
    (1, 3, 16, 24, 64, 32),
    (1, 5, 16, 8, 32, 64),
    (2, 2, 16, 16, 64, 128),
    (1, 2, 16, 8, 32, 256),
    (1, 11, 16, 8, 64, 64),
    (1, 2, 8, 8, 32, 32),          # H < 16: rows past the volume are masked in the epilogue
    (2, 3, 24, 16, 32, 64),        # H % 16 != 0
    (1, 2, 8, 8, 256, 128),        # deep-level shape (bottleneck-like): split-K x8
    (1, 4, 16, 16, 160, 256),      # split-K over 5 chunks, two channel passes
    (2, 8, 32, 32, 128, 128),      # level-3 layer of the U-Net at batch 2: split-K x2 over 64 tiles
]


def _poll():
    import ctypes
    from repmode_b200 import lib as L
    code = ctypes.c_int32(0)
    L.check(L.load().mode_poll_error(ctypes.byref(code)), "mode_poll_error")
    assert code.value == 0, f"tcgen05 pipeline timeout code {code.value}"


@pytest.mark.parametrize("shape", SHAPES)
def test_conv3d_umma_bitexact(shape):
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, k, nout = shape
    rng = np.random.RandomState(sum(shape))
    x = rng.randint(-3, 4, size=(n, d, h, w, k)).astype(np.float32)
    weff = (rng.randint(-4, 5, size=(n, nout, k, 5, 5, 5)) / 8.0).astype(np.float32)
    su = torch.arange(n, dtype=torch.int32, device="cuda")
    xg = torch.from_numpy(x).cuda()
    w32 = torch.from_numpy(pack_weights(weff, half=False)).cuda()
    w16 = torch.from_numpy(pack_weights(weff, half=True)).cuda()
    sums32 = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    sums16 = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    y_simt = Fm.conv3d(xg, L.MODE_F32, w32, su, n, d, h, w, k, nout, None, sums32, impl=L.IMPL_SIMT)
    y_umma = Fm.conv3d(xg.half(), L.MODE_F16, w16, su, n, d, h, w, k, nout, None, sums16, impl=L.IMPL_UMMA)
    _poll()
    if n * d * h * w * k * nout <= 2 ** 22:      # oracle finishes in seconds at these sizes
        ref = np.stack([onp.conv3d_fwd(x[i].transpose(3, 0, 1, 2), weff[i]) for i in range(n)])
        assert np.array_equal(y_simt.cpu().numpy(), ref.transpose(0, 2, 3, 4, 1))
    assert torch.equal(y_umma, y_simt)
    # the sums pass through fp32 per-warp partials (32 voxels) before the fp64 accumulation: exact while |y|^2 * 32 stays
    # below 2^24 / 64 (small K), rounded differently by the two kernels beyond that
    assert torch.allclose(sums16, sums32, rtol=1e-12 if k <= 64 else 1e-6, atol=1e-9)


@pytest.mark.parametrize("cap", [1, 3])
def test_conv3d_umma_splitk_cap_bitexact(cap, monkeypatch):
    """The K split of the single-CTA kernel (deep small-volume layers) is a launch-plan choice: unsplit (cap 1), an uneven
    split (5 chunks over 3 CTAs: 2 + 2 + 1) and the default plan give the same bits on integer-valued data, including the
    BatchNorm sums over a plane sub-range, and the workspace query follows the cap."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, k, nout = 1, 4, 16, 16, 160, 64
    rng = np.random.RandomState(11)
    x = rng.randint(-3, 4, size=(n, d, h, w, k)).astype(np.float32)
    weff = (rng.randint(-4, 5, size=(n, nout, k, 5, 5, 5)) / 8.0).astype(np.float32)
    su = torch.zeros(n, dtype=torch.int32, device="cuda")
    xg = torch.from_numpy(x).cuda()
    w32 = torch.from_numpy(pack_weights(weff, half=False)).cuda()
    w16 = torch.from_numpy(pack_weights(weff, half=True)).cuda()
    sums32 = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    y_simt = Fm.conv3d(xg, L.MODE_F32, w32, su, n, d, h, w, k, nout, None, sums32, impl=L.IMPL_SIMT, stat_range=(1, 3))
    monkeypatch.setenv("REPMODE_UMMA_SPLITK", str(cap))
    lib = L.load()
    ws = int(lib.mode_conv3d_workspace_bytes(n, d, h, w, k, nout, L.MODE_F16))
    assert ws == (0 if cap == 1 else 3 * n * d * h * w * nout * 4)
    sums16 = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    y = Fm.conv3d(xg.half(), L.MODE_F16, w16, su, n, d, h, w, k, nout, None, sums16, impl=L.IMPL_UMMA, stat_range=(1, 3))
    _poll()
    assert torch.equal(y, y_simt)
    assert torch.allclose(sums16, sums32, rtol=1e-6, atol=1e-9)      # fp32 per-thread partials (see above)


PAIR_SHAPES = [  # N, D, H, W, K, Nout, forced clusters (0 = the launcher's choice)
    (1, 1, 16, 8, 32, 32, 0),          # H < 32: the peer CTA's rows are all masked
    (2, 4, 32, 8, 32, 32, 0),
    (1, 9, 32, 16, 32, 32, 0),         # runs cut inside a column: halo planes + garbage slots on both sides
    (1, 30, 32, 8, 32, 32, 1),         # one cluster marches 30 planes: slot period wraps, spill slots 12..15
    (2, 17, 32, 8, 32, 32, 1),         # two columns back to back on one cluster (accumulator hand-off across runs)
    (1, 30, 32, 8, 32, 32, 3),
    (1, 3, 40, 24, 64, 32, 0),         # two K chunks, H % 32 != 0
    (1, 5, 32, 8, 32, 64, 0),          # two channel passes
    (2, 13, 64, 16, 64, 64, 0),
    (1, 2, 8, 8, 128, 96, 2),
    (3, 6, 32, 8, 32, 32, 1),          # one cluster walks three samples: the resident weights are reloaded twice
    (2, 7, 32, 16, 32, 64, 2),
]


@pytest.mark.parametrize("shape", PAIR_SHAPES)
def test_conv3d_pair_bitexact(shape, monkeypatch):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2, impl=4) == SIMT fp32 kernel, exactly, incl. the fused BN sums
    over a plane sub-range."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, k, nout, clusters = shape
    if clusters:
        monkeypatch.setenv("REPMODE_PAIR_CLUSTERS", str(clusters))
    rng = np.random.RandomState(sum(shape))
    x = rng.randint(-3, 4, size=(n, d, h, w, k)).astype(np.float32)
    weff = (rng.randint(-4, 5, size=(n, nout, k, 5, 5, 5)) / 8.0).astype(np.float32)
    su = torch.arange(n, dtype=torch.int32, device="cuda")
    xg = torch.from_numpy(x).cuda()
    w32 = torch.from_numpy(pack_weights(weff, half=False)).cuda()
    w16 = torch.from_numpy(pack_weights(weff, half=True)).cuda()
    sums32 = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    sums16 = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    rng_stat = (1, max(2, d - 1)) if d > 2 else None
    y_simt = Fm.conv3d(xg, L.MODE_F32, w32, su, n, d, h, w, k, nout, None, sums32, impl=L.IMPL_SIMT, stat_range=rng_stat)
    y_pair = Fm.conv3d(xg.half(), L.MODE_F16, w16, su, n, d, h, w, k, nout, None, sums16, impl=L.IMPL_UMMA_PAIR,
                       stat_range=rng_stat)
    _poll()
    assert torch.equal(y_pair, y_simt)
    assert torch.allclose(sums16, sums32, rtol=1e-12, atol=1e-9)


def test_conv3d_umma_random_vs_simt():
    """Real-valued operands: fp16 operand rounding only (<= 2^-11 relative per operand)."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, k, nout = 2, 6, 32, 16, 64, 64
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(n, d, h, w, k, generator=g).cuda()
    weff = (torch.randn(n, nout, k, 5, 5, 5, generator=g) * 0.02).numpy()
    su = torch.arange(n, dtype=torch.int32, device="cuda")
    w32 = torch.from_numpy(pack_weights(weff, half=False)).cuda()
    w16 = torch.from_numpy(pack_weights(weff * 256.0, half=True)).cuda()
    inv = torch.tensor([1.0 / 256.0], device="cuda")
    y_simt = Fm.conv3d(x, L.MODE_F32, w32, su, n, d, h, w, k, nout, None, None, impl=L.IMPL_SIMT)
    y_umma = Fm.conv3d(x.half(), L.MODE_F16, w16, su, n, d, h, w, k, nout, inv, None, impl=L.IMPL_UMMA)
    _poll()
    assert_close(y_umma.cpu().numpy(), y_simt.cpu().numpy(), 1e-3, "umma vs simt")


def test_conv3d_dgrad_pack_matches_oracle():
    """K3 = K2 on the flipped / io-transposed pack: bit-exact against the oracle's dgrad."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, ci, co = 1, 3, 16, 8, 32, 64
    rng = np.random.RandomState(7)
    dy = rng.randint(-3, 4, size=(n, d, h, w, co)).astype(np.float32)
    weff = (rng.randint(-4, 5, size=(n, co, ci, 5, 5, 5)) / 8.0).astype(np.float32)
    su = torch.zeros(n, dtype=torch.int32, device="cuda")
    wd16 = torch.from_numpy(pack_weights(weff, half=True, dgrad=True)).cuda()
    wd32 = torch.from_numpy(pack_weights(weff, half=False, dgrad=True)).cuda()
    dyg = torch.from_numpy(dy).cuda()
    dx_simt = Fm.conv3d(dyg, L.MODE_F32, wd32, su, n, d, h, w, co, ci, None, None, impl=L.IMPL_SIMT)
    dx_umma = Fm.conv3d(dyg.half(), L.MODE_F16, wd16, su, n, d, h, w, co, ci, None, None, impl=L.IMPL_UMMA)
    _poll()
    ref = onp.conv3d_dgrad(dy[0].transpose(3, 0, 1, 2), weff[0]).transpose(1, 2, 3, 0)[None]
    assert np.array_equal(dx_simt.cpu().numpy(), ref)
    assert torch.equal(dx_umma, dx_simt)


WG_SHAPES = [  # N, D, H, W, Ci, Co
    (1, 1, 16, 8, 32, 32),
    (1, 3, 16, 8, 32, 32),
    (2, 4, 32, 16, 32, 32),
    (1, 5, 24, 8, 64, 32),
    (1, 2, 16, 16, 32, 64),
    (1, 6, 40, 24, 64, 64),
    (1, 7, 8, 8, 32, 32),           # H < 16, odd D (the leftover units' last plane pair holds one x plane)
    (1, 9, 48, 16, 32, 32),         # many slabs per unit kind: cost-weighted slab cuts inside planes
]


@pytest.mark.parametrize("impl", ["split", "deep"])
@pytest.mark.parametrize("shape", WG_SHAPES)
def test_wgrad_umma_bitexact(shape, impl):
    """K4 on tcgen05 (tap-stacking through overlapping MN-major views; wgrad_deep.cu = the default, wgrad_split.cu = the
    round-1 kernel kept as the A/B arm) == SIMT fp32 wgrad == oracle, exactly, on integer-valued data."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, ci, co = shape
    which = {"split": L.IMPL_WGRAD_SPLIT, "deep": L.IMPL_WGRAD_DEEP}[impl]
    rng = np.random.RandomState(sum(shape) + 1)
    x = rng.randint(-3, 4, size=(n, d, h, w, ci)).astype(np.float32)
    dy = rng.randint(-3, 4, size=(n, d, h, w, co)).astype(np.float32)
    xg, dyg = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    dw_simt = Fm.conv3d_wgrad(xg, dyg, L.MODE_F32, n, d, h, w, ci, co, None, impl=L.IMPL_SIMT)
    dw_umma = Fm.conv3d_wgrad(xg.half(), dyg.half(), L.MODE_F16, n, d, h, w, ci, co, None, impl=which)
    _poll()
    if n * d * h * w * ci * co <= 2 ** 22:
        ref = np.stack([onp.conv3d_wgrad(x[i].transpose(3, 0, 1, 2), dy[i].transpose(3, 0, 1, 2)) for i in range(n)])
        ref = ref.reshape(n, co, ci, 125).transpose(0, 3, 1, 2)            # [n][tap][o][i]
        assert np.array_equal(dw_simt.cpu().numpy(), ref)
    assert torch.equal(dw_umma, dw_simt)


def test_headline_layer_full_size_bitexact_and_adjoint():
    """BASELINE.json's metric shape -- 32 -> 32 channels on a 1x32x128x128 volume -- through exactly the kernels bench.py
    times (the launcher's own choices: CTA-pair conv for forward and dgrad, the default tcgen05 wgrad with its one-wave
    slab plan): bit-identical to the fp32 SIMT kernels on integer-valued data, and the size-independent adjoint
    identities of a convolution hold EXACTLY between the three kernels:
        <conv(x, W), r> = <x, dgrad(r, W)> = <wgrad(x, r), W>."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, h, w, ci, co = 1, 32, 128, 128, 32, 32
    rng = np.random.RandomState(2024)
    x = rng.randint(-3, 4, size=(n, d, h, w, ci)).astype(np.float32)
    r = rng.randint(-3, 4, size=(n, d, h, w, co)).astype(np.float32)
    weff = (rng.randint(-4, 5, size=(n, co, ci, 5, 5, 5)) / 8.0).astype(np.float32)
    su = torch.zeros(n, dtype=torch.int32, device="cuda")
    xg, rg = torch.from_numpy(x).cuda(), torch.from_numpy(r).cuda()
    x16, r16 = xg.half(), rg.half()

    # forward: |y| <= 4000 * 1.5 in steps of 1/8 -> every partial sum is exact in fp32 in any order
    w32 = torch.from_numpy(pack_weights(weff, half=False)).cuda()
    w16 = torch.from_numpy(pack_weights(weff, half=True)).cuda()
    y_simt = Fm.conv3d(xg, L.MODE_F32, w32, su, n, d, h, w, ci, co, None, None, impl=L.IMPL_SIMT)
    y_auto = Fm.conv3d(x16, L.MODE_F16, w16, su, n, d, h, w, ci, co, None, None, impl=L.IMPL_UMMA)
    _poll()
    assert torch.equal(y_auto, y_simt)

    # dgrad = the same kernel on the flipped / transposed pack
    wd32 = torch.from_numpy(pack_weights(weff, half=False, dgrad=True)).cuda()
    wd16 = torch.from_numpy(pack_weights(weff, half=True, dgrad=True)).cuda()
    dx_simt = Fm.conv3d(rg, L.MODE_F32, wd32, su, n, d, h, w, co, ci, None, None, impl=L.IMPL_SIMT)
    dx_auto = Fm.conv3d(r16, L.MODE_F16, wd16, su, n, d, h, w, co, ci, None, None, impl=L.IMPL_UMMA)
    _poll()
    assert torch.equal(dx_auto, dx_simt)

    # wgrad: |sum| <= 9 * 524288 < 2^24 -> exact
    dw_simt = Fm.conv3d_wgrad(xg, rg, L.MODE_F32, n, d, h, w, ci, co, None, impl=L.IMPL_SIMT)
    dw_auto = Fm.conv3d_wgrad(x16, r16, L.MODE_F16, n, d, h, w, ci, co, None)
    _poll()
    assert torch.equal(dw_auto, dw_simt)

    # adjoint identities in fp64 (all terms are multiples of 1/8 far below 2^53: exact)
    a = float((y_auto.double() * rg.double()).sum())
    b = float((xg.double() * dx_auto.double()).sum())
    w_tap = torch.from_numpy(weff.reshape(n, co, ci, 125).transpose(0, 3, 1, 2).copy()).cuda()    # [n][tap][o][i]
    c = float((dw_auto.double() * w_tap.double()).sum())
    assert a == b == c, (a, b, c)


EX_SHAPES = [  # N, D (output planes), lo halo, hi halo, H, W, K, Nout, impl
    (1, 6, 2, 2, 32, 16, 32, 32, "pair"),
    (1, 9, 2, 0, 32, 8, 32, 32, "pair"),        # slab at the upper global face: no halo above
    (2, 5, 0, 2, 32, 16, 32, 64, "pair"),
    (1, 12, 4, 4, 32, 8, 64, 32, "pair"),       # streaming plan, 4-plane halo (two-conv stage, conv1)
    (1, 6, 2, 2, 16, 8, 32, 32, "single"),
    (2, 7, 2, 3, 24, 16, 64, 64, "single"),
    (1, 3, 4, 4, 16, 8, 32, 128, "single"),
    (1, 2, 1, 1, 8, 8, 96, 64, "single"),          # split-K x3 with halo, epilogue and fp16-only result
    (1, 4, 0, 2, 16, 16, 256, 128, "single"),
    (1, 4, 2, 2, 12, 10, 5, 7, "simt"),
]


@pytest.mark.parametrize("shape", EX_SHAPES)
def test_conv3d_ex_haloed_input_and_epilogue_bitexact(shape):
    """mode_conv3d_ex: a haloed input (Dx = lo + D + hi planes, output plane q centred on input plane q + lo) gives exactly the
    owned planes of the plain 'same' conv over the extended slab, on all three kernels; the fused epilogue (per-channel
    affine + ReLU, fp16 copy into a plane window of a larger buffer) matches the same arithmetic done afterwards."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, lo, hi, h, w, k, nout, which = shape
    dx = lo + d + hi
    rng = np.random.RandomState(sum(shape[:8]))
    x = rng.randint(-3, 4, size=(n, dx, h, w, k)).astype(np.float32)
    weff = (rng.randint(-4, 5, size=(n, nout, k, 5, 5, 5)) / 8.0).astype(np.float32)
    su = torch.arange(n, dtype=torch.int32, device="cuda")
    xg = torch.from_numpy(x).cuda()
    w32 = torch.from_numpy(pack_weights(weff, half=False)).cuda()
    ref_full = Fm.conv3d(xg, L.MODE_F32, w32, su, n, dx, h, w, k, nout, None, None, impl=L.IMPL_SIMT)
    ref = ref_full[:, lo:lo + d].contiguous()
    if which == "simt":
        xin, win, dt, impl = xg, w32, L.MODE_F32, L.IMPL_SIMT
    else:
        xin, win, dt = xg.half(), torch.from_numpy(pack_weights(weff, half=True)).cuda(), L.MODE_F16
        impl = L.IMPL_UMMA_PAIR if which == "pair" else L.IMPL_UMMA_SINGLE
    sums = torch.zeros(2 * nout, dtype=torch.float64, device="cuda")
    y = Fm.conv3d(xin, dt, win, su, n, d, h, w, k, nout, None, sums, impl=impl, halo=(dx, lo))
    _poll()
    assert torch.equal(y, ref)
    assert torch.allclose(sums[:nout], ref.double().sum(dim=(0, 1, 2, 3)), rtol=1e-12, atol=1e-9)
    # fused epilogue: power-of-two scale, integer shift -> exact in fp32 either way
    sc = torch.from_numpy(2.0 ** rng.randint(-2, 2, size=nout)).float().cuda()
    sh = torch.from_numpy(rng.randint(-5, 6, size=nout).astype(np.float32)).cuda()
    y16 = torch.full((n, d + 3, h, w, nout), 7.0, dtype=torch.float16, device="cuda")
    y2 = Fm.conv3d(xin, dt, win, su, n, d, h, w, k, nout, None, None, impl=impl, halo=(dx, lo), ep=(sc, sh, True),
                   y16=(y16, 1, 0.25))
    _poll()
    want = torch.relu(ref * sc + sh)
    assert torch.equal(y2, want)
    assert torch.equal(y16[:, 1:1 + d], (want * 0.25).clamp(-65504, 65504).half())
    assert torch.all(y16[:, 0] == 7.0) and torch.all(y16[:, 1 + d:] == 7.0)       # planes outside the window untouched
    if which == "pair" and k > 32:
        return                        # K > 32 on the pair kernel accumulates chunk by chunk in the fp32 output
    y16b = torch.zeros((n, d, h, w, nout), dtype=torch.float16, device="cuda")
    none = Fm.conv3d(xin, dt, win, su, n, d, h, w, k, nout, None, None, impl=impl, halo=(dx, lo), ep=(sc, sh, True),
                     y16=(y16b, 0, 0.25), want_y=False)
    _poll()
    assert none is None and torch.equal(y16b, y16[:, 1:1 + d])


@pytest.mark.parametrize("shape", [(1, 6, 2, 2, 32, 16, 32, 32), (1, 5, 2, 0, 16, 8, 32, 64), (2, 4, 0, 2, 24, 8, 64, 32),
                                   (1, 8, 4, 4, 16, 8, 32, 32), (1, 3, 3, 1, 8, 8, 32, 32)])
@pytest.mark.parametrize("impl", ["deep", "simt"])
def test_wgrad_ex_haloed_x_bitexact(shape, impl):
    """mode_conv3d_wgrad_ex: dy = the D owned planes, x = lo + D + hi planes == the plain wgrad over the extended slab with
    dy zero outside the owned planes, exactly."""
    from repmode_b200 import functional as Fm, lib as L
    n, d, lo, hi, h, w, ci, co = shape
    dx = lo + d + hi
    rng = np.random.RandomState(sum(shape) + 3)
    x = rng.randint(-3, 4, size=(n, dx, h, w, ci)).astype(np.float32)
    dy = rng.randint(-3, 4, size=(n, d, h, w, co)).astype(np.float32)
    dy_ext = np.zeros((n, dx, h, w, co), dtype=np.float32)
    dy_ext[:, lo:lo + d] = dy
    xg, dyg, dyeg = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda(), torch.from_numpy(dy_ext).cuda()
    ref = Fm.conv3d_wgrad(xg, dyeg, L.MODE_F32, n, dx, h, w, ci, co, None, impl=L.IMPL_SIMT)
    if impl == "simt":
        got = Fm.conv3d_wgrad(xg, dyg, L.MODE_F32, n, d, h, w, ci, co, None, impl=L.IMPL_SIMT, halo=(dx, lo))
    else:
        got = Fm.conv3d_wgrad(xg.half(), dyg.half(), L.MODE_F16, n, d, h, w, ci, co, None, impl=L.IMPL_WGRAD_DEEP, halo=(dx, lo))
    _poll()
    assert torch.equal(got, ref)
