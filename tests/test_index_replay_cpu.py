"""CPU replays of index arithmetic that two CUDA translation units (or a kernel and its host caller) must agree on -- the
contracts a parity test on the GPU would only report as "numbers differ":

 * K1b reading d_weff straight out of K4's work-unit partials (reparam_bwd_reg_kernel<true>, repmode_b200/csrc/reparam.cu)
   against the unit / entry layout the wgrad kernel writes and its reduce kernel reads (wgrad_deep_reduce_kernel,
   repmode_b200/csrc/wgrad_deep.cu);
 * the depth-to-space row map of the BatchNorm kernels (RowMap::offset, repmode_b200/csrc/bn.cu) against the permute +
   reshape it replaces (ConvTranspose3d(k=2, s=2) as a GEMM, fnet/nn_modules/RepMode.py:97-101);
 * the block -> (layer, row block, chunk) lookup of the grouped K1 launch (reparam_fwd_rows_grouped_kernel) and the tile ->
   layer lookup of the grouped dgrad pack, including layers without a dgrad pack (empty tile range)."""
import numpy as np
import pytest
import torch

PF = 60 * 32 * 32           # K4_DEEP_PARTIAL_FLOATS (common.cuh)


def reduce_source(n, tap, o, i, ncic, ncoc, SL, SA, SB, nL, nA):
    """(first unit, slabs, float offset inside a unit) of d_weff[n, tap, o, i]: wgrad_deep_reduce_kernel."""
    kd, kh, kw = tap // 25, (tap // 5) % 5, tap % 5
    g = (n * ncoc + (o >> 5)) * ncic + (i >> 5)
    if kh == 4:
        unit0, S, entry = g * SL, SL, kd * 5 + kw
    elif kd < 2:
        unit0, S, entry = nL + g * SA, SA, kd * 20 + kh * 5 + kw
    else:
        unit0, S, entry = nL + nA + g * SB, SB, (kd - 2) * 20 + kh * 5 + kw
    return unit0, S, (entry * 32 + (o & 31)) * 32 + (i & 31)


def k1b_partial_offset(n, tap, o, ic, cq, ncic, ncoc, nL, nA):
    """Float offset of the 16 bytes thread (tap, channel quad cq) of block (ic, o) loads for sample n:
    reparam_bwd_reg_kernel<true> (poff + n * pstride)."""
    kd, kh, kw = tap // 25, (tap // 5) % 5, tap % 5
    g0 = (o >> 5) * ncic + ic
    if kh == 4:
        unit, entry = g0, kd * 5 + kw
    elif kd < 2:
        unit, entry = nL + g0, kd * 20 + kh * 5 + kw
    else:
        unit, entry = nL + nA + g0, (kd - 2) * 20 + kh * 5 + kw
    poff = unit * PF + (entry * 32 + (o & 31)) * 32 + cq * 4
    return poff + n * (ncoc * ncic * PF)


@pytest.mark.parametrize("N,ci,co", [(4, 128, 128), (2, 256, 160), (1, 512, 512)])
def test_k1b_reads_the_partials_where_the_reduce_kernel_reads_them(N, ci, co):
    ncic, ncoc = ci // 32, co // 32
    groups = N * ncic * ncoc
    SL = SA = SB = 1                       # the only plan mode_reparam_bwd_partial accepts
    nL, nA = groups * SL, groups * SA
    rng = np.random.default_rng(0)
    seen = set()
    for _ in range(4000):
        n, tap, o, i = int(rng.integers(N)), int(rng.integers(125)), int(rng.integers(co)), int(rng.integers(ci))
        unit0, S, off = reduce_source(n, tap, o, i, ncic, ncoc, SL, SA, SB, nL, nA)
        assert S == 1
        want = unit0 * PF + off
        got = k1b_partial_offset(n, tap, o, i >> 5, (i & 31) >> 2, ncic, ncoc, nL, nA) + (i & 3)
        assert got == want, (n, tap, o, i)
        assert want < 3 * groups * PF
        seen.add(want)
    assert len(seen) > 3900                 # distinct elements map to distinct addresses


def d2s_row(row, D, H, W):
    """RowMap::offset / pitch for d2s: row of the [voxels][8] GEMM result -> row of the NDHWC volume of twice the size."""
    tap, m = row & 7, row >> 3
    w, m = m % W, m // W
    h, m = m % H, m // H
    d, n = m % D, m // D
    return ((n * 2 * D + 2 * d + (tap >> 2)) * 2 * H + 2 * h + ((tap >> 1) & 1)) * 2 * W + 2 * w + (tap & 1)


@pytest.mark.parametrize("n,d,h,w,co", [(2, 3, 4, 5, 4), (1, 1, 2, 8, 8)])
def test_depth_to_space_row_map_equals_the_permute_it_replaces(n, d, h, w, co):
    rows = n * d * h * w * 8
    y8 = torch.arange(rows * co, dtype=torch.int64).view(n * d * h * w, 8 * co)         # GEMM result [voxels, (kd,kh,kw,co)]
    want = y8.view(n, d, h, w, 2, 2, 2, co).permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(n, 2 * d, 2 * h, 2 * w, co)
    got = torch.full((rows, co), -1, dtype=torch.int64)
    src = y8.view(rows, co)
    for r in range(rows):
        got[d2s_row(r, d, h, w)] = src[r]
    assert torch.equal(got.view(n, 2 * d, 2 * h, 2 * w, co), want)
    assert sorted(d2s_row(r, d, h, w) for r in range(rows)) == list(range(rows))            # a permutation of the rows


def test_grouped_k1_block_and_tile_lookup():
    layers = [(32, 32, True), (32, 64, True), (64, 64, False), (128, 32, True), (96, 160, True)]   # (ci, co, has dgrad pack)
    U, ROWS = 3, 2
    block_begin, tile_begin, blocks, tiles = [], [], 0, 0
    for ci, co, dg in layers:                                   # mode_reparam_fwd_grouped's host loop
        block_begin.append(blocks)
        tile_begin.append(tiles)
        blocks += (co // ROWS) * (ci // 32)
        if dg:
            tiles += U * 125 * (ci // 32) * (co // 32)

    def locate(begin, b):                                       # the kernels' scan
        k = 0
        while k + 1 < len(layers) and b >= begin[k + 1]:
            k += 1
        return k
    covered = [set() for _ in layers]
    for b in range(blocks):
        k = locate(block_begin, b)
        ci, co, _ = layers[k]
        local = b - block_begin[k]
        row_blocks = co // ROWS
        rb, chunk = local % row_blocks, local // row_blocks
        assert 0 <= chunk < ci // 32
        covered[k].add((rb, chunk))
    for k, (ci, co, _) in enumerate(layers):
        assert len(covered[k]) == (co // ROWS) * (ci // 32)     # every (row block, chunk) of every layer exactly once
    per_layer = [0] * len(layers)
    for t in range(0, tiles, 37):
        k = locate(tile_begin, t)
        assert layers[k][2], "a tile landed in a layer without a dgrad pack"
        assert 0 <= t - tile_begin[k] < U * 125 * (layers[k][0] // 32) * (layers[k][1] // 32)
        per_layer[k] += 1
    assert per_layer[2] == 0 and all(c > 0 for k, c in enumerate(per_layer) if k != 2)
