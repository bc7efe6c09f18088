"""CPU (-m "not gpu"): the TAP COVER of the tcgen05 wgrad kernels, replayed in numpy.

wgrad_split.cu / wgrad_deep.cu compute d_weff by stacking taps through shifted operand views: an MMA
multiplies a dy brick seen through 4 "M blocks" (row shifts, or plane shifts for the kh = 4 units) with an x brick seen
through 5 "N blocks" (voxel shifts), over K steps of 2 rows x 8 voxels, and the epilogue maps (accumulator set, M
block, N block) to a tap.  This file restates each kernel's SCHEDULE -- which bricks a tile loads (zero fill outside
the volume), which MMAs it issues, where the epilogue files every accumulator block -- with plain loops, and checks
that the sum is exactly the oracle's wgrad: every tap produced once, padding right, no plane or row lost by the K-step
trimming or by the plane-pair tiling.  It checks the index algebra only (the descriptor / swizzle semantics are what the
-m gpu bit-exact tests cover).
"""
import numpy as np
import pytest

from oracle import mode_numpy as onp

TW, TH = 8, 16


def _pad_get(a, d, h, w):
    """a[d, h, w, :] with zero fill outside the volume (TMA out-of-bounds fill)."""
    D, H, W, _ = a.shape
    if 0 <= d < D and 0 <= h < H and 0 <= w < W:
        return a[d, h, w]
    return np.zeros(a.shape[-1], a.dtype)


def _mma(acc, dy, x, dy_org, m_shifts, x_org, kk):
    """One M128 N160 K16 MMA: acc[bm][bn] += sum over the K step's 2 rows x 8 voxels of dy(view bm)^T x(view bn).
    dy_org / x_org = (plane, row, col) of the brick row the descriptors start at; m_shifts[bm] = (dplane, drow)."""
    for bm, (dd, dh) in enumerate(m_shifts):
        for bn in range(5):
            for r in range(2):
                for c in range(TW):
                    a = _pad_get(dy, dy_org[0] + dd, dy_org[1] + 2 * kk + r + dh, dy_org[2] + c)
                    b = _pad_get(x, x_org[0], x_org[1] + 2 * kk + r, x_org[2] + c + bn)
                    acc[bm][bn] += np.outer(a, b)


ROW_SHIFTS = [(0, bm) for bm in range(4)]       # M block bm = dy brick + bm rows      (kh = 3 - bm)
PLANE_SHIFTS = [(bm, 0) for bm in range(4)]     # M block bm = dy plane + bm           (kd = 3 - bm / 4 - bm)


def _new_acc(co, ci):
    return [[np.zeros((co, ci)) for _ in range(5)] for _ in range(4)]


def _tiles_h(H, offset):
    return -(-(H + offset) // TH)


def _l_units(x, dy, dw, always_two_x_planes):
    """kh = 4 of all kd: 6-plane dy box p-2..p+3, x planes p, p+1 with rows +2; M blocks = plane shifts."""
    D, H, W, ci = x.shape
    co = dy.shape[-1]
    set0, set1 = _new_acc(co, ci), _new_acc(co, ci)
    th_n = _tiles_h(H, 0)
    for tg in range((D + 1) // 2):
        p = 2 * tg
        nx = 2 if (always_two_x_planes or p + 1 < D) else 1
        for th in range(th_n):
            vh0 = th * TH
            nk = min(8, (H - vh0 + 1) >> 1)
            for tw in range(W // TW):
                vw0 = tw * TW
                for kk in range(nk):
                    for xi in range(nx):
                        xo = (p + xi, vh0 + 2, vw0 - 2)
                        _mma(set0, dy, x, (p - 2 + 1 + xi, vh0, vw0), PLANE_SHIFTS, xo, kk)   # planes p-1+xi .. : kd = 3 - bm
                        _mma(set1, dy, x, (p - 2 + xi, vh0, vw0), PLANE_SHIFTS, xo, kk)       # planes p-2+xi .. : bm = 0 is kd = 4
    for bm in range(4):
        for bn in range(5):
            dw[(3 - bm) * 25 + 4 * 5 + bn] += set0[bm][bn]
            if bm == 0:
                dw[4 * 25 + 4 * 5 + bn] += set1[bm][bn]


def cover_split(x, dy):
    """wgrad_split.cu: K units (kh = 0..3 of kd pairs (0,1), (2,3) and of kd = 4; x planes outside the volume skipped)
    + L units."""
    D, H, W, ci = x.shape
    co = dy.shape[-1]
    dw = np.zeros((125, co, ci))
    th_n = _tiles_h(H, 3)
    for kd0, nkd in ((0, 2), (2, 2), (4, 1)):
        sets = [_new_acc(co, ci) for _ in range(nkd)]
        dlo, dhi = max(0, 2 - (kd0 + nkd - 1)), min(D, D + 2 - kd0)
        for d in range(dlo, dhi):
            for th in range(th_n):
                vh0 = th * TH - 3
                nk = min(8, (H - vh0 + 1) >> 1)
                for tw in range(W // TW):
                    vw0 = tw * TW
                    for j in range(nkd):
                        xp = d + kd0 + j - 2
                        if 0 <= xp < D:
                            for kk in range(nk):
                                _mma(sets[j], dy, x, (d, vh0, vw0), ROW_SHIFTS, (xp, vh0 + 1, vw0 - 2), kk)
        for j in range(nkd):
            for bm in range(4):
                for bn in range(5):
                    dw[(kd0 + j) * 25 + (3 - bm) * 5 + bn] += sets[j][bm][bn]
    _l_units(x, dy, dw, always_two_x_planes=False)
    return dw


def cover_deep(x, dy, x_off=0):
    """wgrad_deep.cu: A units (kd 0,1) and B units (kd 2,3,4) walk dy planes in pairs (one 2-plane dy box, one x box of
    the 3 / 4 planes they meet; planes outside the volume are zero fill and multiplied like any other) + L units.
    Haloed form (mode_conv3d_wgrad_ex): x has Dx >= D planes and dy plane p is centred on x plane p + x_off."""
    Dx, H, W, ci = x.shape
    D = dy.shape[0]
    co = dy.shape[-1]
    dw = np.zeros((125, co, ci))
    th_n = _tiles_h(H, 3)
    PT = 2
    for kd0, nkd in ((0, 2), (2, 3)):
        sets = [_new_acc(co, ci) for _ in range(nkd)]
        dlo, dhi = max(0, 2 - (kd0 + nkd - 1) - x_off), min(D, Dx + 2 - x_off - kd0)
        for tg in range((max(0, dhi - dlo) + PT - 1) // PT):
            d = dlo + PT * tg
            for th in range(th_n):
                vh0 = th * TH - 3
                nk = min(8, (H - vh0 + 1) >> 1)
                for tw in range(W // TW):
                    vw0 = tw * TW
                    for pl in range(PT):
                        for kk in range(nk):
                            for j in range(nkd):
                                # x box starts at x plane d + kd0 - 2 + x_off; local plane pl + j
                                _mma(sets[j], dy, x, (d + pl, vh0, vw0), ROW_SHIFTS,
                                     (d + kd0 - 2 + x_off + pl + j, vh0 + 1, vw0 - 2), kk)
        for j in range(nkd):
            for bm in range(4):
                for bn in range(5):
                    dw[(kd0 + j) * 25 + (3 - bm) * 5 + bn] += sets[j][bm][bn]
    # L units: kh = 4 of all kd; pairs of x planes p, p+1 (dy coordinates) from max(-x_off, -2) to min(Dx-x_off-1, D+1)
    set0, set1 = _new_acc(co, ci), _new_acc(co, ci)
    th_n = _tiles_h(H, 0)
    plo, phi = max(-x_off, -2), min(Dx - x_off - 1, D + 1)
    for tg in range((max(0, phi - plo + 1) + 1) // 2):
        p = plo + 2 * tg
        for th in range(th_n):
            vh0 = th * TH
            nk = min(8, (H - vh0 + 1) >> 1)
            for tw in range(W // TW):
                vw0 = tw * TW
                for kk in range(nk):
                    for xi in range(2):
                        xo = (p + xi + x_off, vh0 + 2, vw0 - 2)
                        _mma(set0, dy, x, (p - 2 + 1 + xi, vh0, vw0), PLANE_SHIFTS, xo, kk)   # planes p-1+xi .. : kd = 3 - bm
                        _mma(set1, dy, x, (p - 2 + xi, vh0, vw0), PLANE_SHIFTS, xo, kk)       # planes p-2+xi .. : bm = 0 is kd = 4
    for bm in range(4):
        for bn in range(5):
            dw[(3 - bm) * 25 + 4 * 5 + bn] += set0[bm][bn]
            if bm == 0:
                dw[4 * 25 + 4 * 5 + bn] += set1[bm][bn]
    return dw


SHAPES = [(1, 16, 8), (3, 16, 8), (4, 24, 16), (5, 8, 8), (2, 40, 8)]      # D, H, W


@pytest.mark.parametrize("cover", [cover_split, cover_deep])
@pytest.mark.parametrize("shape", SHAPES)
def test_tap_cover_equals_oracle_wgrad(cover, shape):
    D, H, W = shape
    ci, co = 2, 3
    rng = np.random.RandomState(D * 100 + H + W)
    x = rng.randint(-3, 4, size=(D, H, W, ci)).astype(np.float64)
    dy = rng.randint(-3, 4, size=(D, H, W, co)).astype(np.float64)
    ref = onp.conv3d_wgrad(x.transpose(3, 0, 1, 2).astype(np.float32), dy.transpose(3, 0, 1, 2).astype(np.float32))
    ref = np.asarray(ref, dtype=np.float64).reshape(co, ci, 125).transpose(2, 0, 1)          # [tap][o][i]
    got = cover(x, dy)
    assert np.array_equal(got, ref), f"{cover.__name__}: {np.argwhere(got != ref)[:5].tolist()}"


@pytest.mark.parametrize("shape", [(3, 7, 2, 16, 8), (4, 8, 2, 8, 8), (2, 10, 4, 16, 8), (5, 7, 0, 16, 8), (3, 6, 3, 8, 8)])
def test_deep_cover_haloed_x(shape):
    """mode_conv3d_wgrad_ex: dy = the D owned planes, x = Dx planes with dy plane p centred on x plane p + x_off."""
    D, Dx, x_off, H, W = shape
    ci, co = 2, 3
    rng = np.random.RandomState(sum(shape))
    x = rng.randint(-3, 4, size=(Dx, H, W, ci)).astype(np.float64)
    dy = rng.randint(-3, 4, size=(D, H, W, co)).astype(np.float64)
    dy_ext = np.zeros((Dx, H, W, co))
    dy_ext[x_off:x_off + D] = dy
    ref = onp.conv3d_wgrad(x.transpose(3, 0, 1, 2).astype(np.float32), dy_ext.transpose(3, 0, 1, 2).astype(np.float32))
    ref = np.asarray(ref, dtype=np.float64).reshape(co, ci, 125).transpose(2, 0, 1)
    got = cover_deep(x, dy, x_off)
    assert np.array_equal(got, ref), np.argwhere(got != ref)[:5].tolist()
