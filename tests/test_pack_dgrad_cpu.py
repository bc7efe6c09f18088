"""CPU replay of the index arithmetic of pack_dgrad_h16_kernel (repmode_b200/csrc/reparam.cu): the vectorised fwd-pack ->
dgrad-pack transposition must place every element where the definition (pack_index of the stage-major, 64-byte-swizzled
fp16 layout: dst[u, 124 - tap, chunk = oc, row = i, col = o] = src[u, tap, chunk = ic, row = o, col = i]) puts it."""
import numpy as np
import pytest


def pack_col(row, col):
    return ((((col >> 3) ^ (row >> 1)) & 3) << 3) | (col & 7)


def pack_index(u, tap, chunk, nchunk, row, nrows, col):
    kd, t = divmod(tap, 25)
    return (((((u * nchunk + chunk) * 25 + t) * 5 + (4 - kd)) * nrows + row) * 32) + pack_col(row, col)


def replay_kernel(src, U, nci, nco, src_rows_pad, dst_rows_pad):
    n_tiles = U * 125 * nci * nco
    dst = np.full(U * 125 * nco * dst_rows_pad * 32, -1, dtype=np.int64)
    for tl in range(n_tiles):
        oc = tl % nco
        q = tl // nco
        ic = q % nci
        q //= nci
        tap, u = q % 125, q // 125
        kd, t = divmod(tap, 25)
        sbase = (((((u * nci + ic) * 25 + t) * 5 + (4 - kd)) * src_rows_pad + oc * 32) * 32)
        dbase = (((((u * nco + oc) * 25 + (24 - t)) * 5 + kd) * dst_rows_pad + ic * 32) * 32)
        tile = np.zeros((32, 34), dtype=np.int64)
        for j in range(128):
            r, pc = j >> 2, j & 3
            lc = pc ^ ((r >> 1) & 3)
            for e in range(8):
                tile[lc * 8 + e][r] = src[sbase + r * 32 + pc * 8 + e]
        for j in range(128):
            r, pc = j >> 2, j & 3
            lc = pc ^ ((r >> 1) & 3)
            for e in range(8):
                dst[dbase + r * 32 + pc * 8 + e] = tile[r][lc * 8 + e]
    return dst


@pytest.mark.parametrize("U,ci,co", [(1, 32, 32), (2, 64, 32), (1, 32, 96)])
def test_pack_dgrad_h16_index_replay(U, ci, co):
    nci, nco = ci // 32, co // 32
    src_rows_pad, dst_rows_pad = nco * 32, nci * 32
    n = U * 125 * nci * src_rows_pad * 32
    src = np.arange(n, dtype=np.int64)                    # every source element carries its own address
    got = replay_kernel(src, U, nci, nco, src_rows_pad, dst_rows_pad)
    want = np.full_like(got, -1)
    for u in range(U):
        for tap in range(0, 125, 7 if ci * co > 1024 else 1):
            for o in range(co):
                for i in range(ci):
                    s = pack_index(u, tap, i // 32, nci, o, src_rows_pad, i % 32)
                    d = pack_index(u, 124 - tap, o // 32, nco, i, dst_rows_pad, o % 32)
                    want[d] = src[s]
    mask = want >= 0
    assert mask.any() and np.array_equal(got[mask], want[mask])
    assert (got >= 0).all()                               # every destination element is written
