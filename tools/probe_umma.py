"""Hardware probe driver (test infrastructure): checks, on a real B200, which shared-memory layouts the
UMMA (tcgen05.mma) and TMA units implement, against numpy models of the hypotheses the conv kernels are
designed around.  Run on the GPU box:  python tools/probe_umma.py  -> gpurun_out/probe_umma.json

Hypotheses
  H1  K-major operands: swizzle is a pure function of the absolute shared-memory address, so the
      descriptor start address may be shifted by whole 64 B / 128 B rows (the conv's kw/kh tap shift)
      and SBO may be any multiple of the row size (brick row pitch) -- with base_offset 0 or "auto".
  H2  MN-major operands: LBO is the stride between swizzle-wide MN blocks and may be as small as one
      row (overlapping, shifted views of the same brick -> tap stacking in wgrad); SBO strides 8-row
      K groups.
  H3  cycles per MMA for the candidate instruction shapes (SS mode).
  H4  TMA tiled 5-d box load: OOB (negative / past-the-end) coordinates zero-fill, swizzle pattern
      equals the address-based model.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libumma_probe.so"))
lib.probe_last_error.restype = ctypes.c_char_p
lib.probe_mma.restype = ctypes.c_int
lib.probe_mma.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                          ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                          ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
lib.probe_tma.restype = ctypes.c_int
lib.probe_tma.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                          ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                          ctypes.c_uint32, ctypes.c_void_p]

SWZ_NONE, SWZ_128B, SWZ_64B, SWZ_32B = 0, 2, 4, 6
FMT_F16, FMT_BF16, FMT_TF32 = 0, 1, 2


def swz(addr, mode):
    addr = np.asarray(addr, dtype=np.int64)
    if mode == SWZ_128B:
        return addr ^ (((addr >> 7) & 7) << 4)
    if mode == SWZ_64B:
        return addr ^ (((addr >> 7) & 3) << 4)
    if mode == SWZ_32B:
        return addr ^ (((addr >> 7) & 1) << 4)
    return addr


def make_desc(addr, lbo, sbo, swizzle, base_offset=0):
    d = (addr >> 4) & 0x3FFF
    d |= ((lbo >> 4) & 0x3FFF) << 16
    d |= ((sbo >> 4) & 0x3FFF) << 32
    d |= 1 << 46
    d |= (base_offset & 7) << 49
    d |= (swizzle & 7) << 61
    return d


def make_idesc(fmt, m, n, a_mn, b_mn):
    d = 1 << 4
    d |= (fmt & 7) << 7
    d |= (fmt & 7) << 10
    d |= (a_mn & 1) << 15
    d |= (b_mn & 1) << 16
    d |= ((n >> 3) & 0x3F) << 17
    d |= ((m >> 4) & 0x1F) << 24
    return d


def tf32_round(x):
    """Round-to-nearest-even fp32 -> tf32 (10 explicit mantissa bits), returned as fp32."""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


class Image:
    """A shared-memory image; writes go through the absolute-address swizzle model (base is 1024-aligned)."""

    def __init__(self, nbytes):
        self.buf = np.zeros(nbytes, dtype=np.uint8)

    def put_rows(self, off, rows, mode):
        """rows: [n_rows, row_bytes] uint8, row r at logical byte address off + r*row_bytes."""
        n, rb = rows.shape
        la = off + (np.arange(n)[:, None] * rb + np.arange(rb)[None, :])
        self.buf[swz(la, mode)] = rows


def run_mma(img, adesc, bdesc, idesc, kind, n_k, a_step, b_step, ncols, repeat=1):
    dev = torch.device("cuda")
    img_t = torch.from_numpy(img.buf).to(dev)
    out = torch.zeros(128 * ncols, dtype=torch.float32, device=dev)
    cyc = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.probe_mma(img_t.data_ptr(), img.buf.size, adesc, bdesc, idesc, kind, n_k, a_step >> 4, b_step >> 4,
                       repeat, ncols, out.data_ptr(), cyc.data_ptr(), st.data_ptr())
    if rc != 0:
        raise RuntimeError(lib.probe_last_error().decode())
    return out.cpu().numpy().reshape(128, ncols), int(cyc.item()), int(st.item())


def relerr(got, ref):
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


RESULTS = []


def record(name, **kw):
    kw["name"] = name
    RESULTS.append(kw)
    print(json.dumps(kw), flush=True)


# ------------------------------------------------------------------------------------------------ H1
def test_kmajor(kind, chans, s, G, bo_mode, N=32, seed=0):
    """A = shifted/strided view of a voxel brick (K-major), B = canonical weights (K-major)."""
    rng = np.random.default_rng(seed)
    esz = 2 if kind == 0 else 4
    rowb = chans * esz                       # 64 or 128
    mode = SWZ_64B if rowb == 64 else SWZ_128B
    kstep_b = 32                             # one MMA consumes 32 bytes of K
    n_k = rowb // kstep_b
    NV = 15 * G + 8 + 8
    if kind == 0:
        X = rng.standard_normal((NV, chans)).astype(np.float16)
        Wt = rng.standard_normal((N, chans)).astype(np.float16)
    else:
        X = tf32_round(rng.standard_normal((NV, chans)).astype(np.float32))
        Wt = tf32_round(rng.standard_normal((N, chans)).astype(np.float32))
    A_OFF = 0
    B_OFF = ((NV * rowb + 1023) // 1024) * 1024
    img = Image(B_OFF + ((N * rowb + 1023) // 1024) * 1024)
    img.put_rows(A_OFF, X.view(np.uint8).reshape(NV, rowb), mode)
    img.put_rows(B_OFF, Wt.view(np.uint8).reshape(N, rowb), mode)
    a_start = A_OFF + s * rowb
    bo = 0 if bo_mode == 0 else ((a_start >> 7) & 7)
    adesc = make_desc(a_start, 16, G * rowb, mode, bo)
    bdesc = make_desc(B_OFF, 16, 8 * rowb, mode, 0)
    idesc = make_idesc(FMT_F16 if kind == 0 else FMT_TF32, 128, N, 0, 0)
    got, cyc, st = run_mma(img, adesc, bdesc, idesc, kind, n_k, kstep_b, kstep_b, N)
    r = np.arange(128)
    vox = (r // 8) * G + (r % 8) + s
    ref = X[vox].astype(np.float64) @ Wt.astype(np.float64).T
    record("H1_kmajor", kind=kind, chans=chans, shift=s, G=G, bo_mode=bo_mode, status=st,
           relerr=relerr(got, ref), ok=bool(relerr(got, ref) < 1e-3))


# ------------------------------------------------------------------------------------------------ H2
def test_mnmajor(sh_a, sh_b, G, swap, NB=5, seed=1):
    """wgrad shape: A = dy^T (M-major, 4 blocks of 32 co, block g shifted by g*sh_a voxels),
    B = x (N-major, NB blocks of 32 ci, block g shifted by g*sh_b voxels); K = 16 voxels = 2 groups of 8
    voxels G voxels apart."""
    rng = np.random.default_rng(seed)
    rowb = 64
    mode = SWZ_64B
    NV = 2 * G + 8 + 8 * max(sh_a, sh_b) + 16
    DY = rng.standard_normal((NV, 32)).astype(np.float16)
    XX = rng.standard_normal((NV, 32)).astype(np.float16)
    A_OFF = 0
    B_OFF = ((NV * rowb + 1023) // 1024) * 1024
    img = Image(2 * B_OFF)
    img.put_rows(A_OFF, DY.view(np.uint8).reshape(NV, rowb), mode)
    img.put_rows(B_OFF, XX.view(np.uint8).reshape(NV, rowb), mode)
    blk_a, blk_b, kgrp = sh_a * rowb, sh_b * rowb, G * rowb
    if not swap:   # LBO = MN-block stride, SBO = K-group stride
        adesc = make_desc(A_OFF, blk_a, kgrp, mode)
        bdesc = make_desc(B_OFF, blk_b, kgrp, mode)
    else:
        adesc = make_desc(A_OFF, kgrp, blk_a, mode)
        bdesc = make_desc(B_OFF, kgrp, blk_b, mode)
    N = 32 * NB
    idesc = make_idesc(FMT_F16, 128, N, 1, 1)
    got, cyc, st = run_mma(img, adesc, bdesc, idesc, 0, 1, 0, 0, N)
    k = np.arange(16)
    kv = (k // 8) * G + (k % 8)
    m = np.arange(128)
    n = np.arange(N)
    Am = DY[(m // 32)[:, None] * sh_a + kv[None, :], (m % 32)[:, None]].astype(np.float64)   # [128,16]
    Bn = XX[(n // 32)[:, None] * sh_b + kv[None, :], (n % 32)[:, None]].astype(np.float64)   # [N,16]
    ref = Am @ Bn.T
    record("H2_mnmajor", sh_a=sh_a, sh_b=sh_b, G=G, swap=swap, NB=NB, status=st, relerr=relerr(got, ref),
           ok=bool(relerr(got, ref) < 1e-3))


# ------------------------------------------------------------------------------------------------ H3
def bench_shape(kind, N, mn_major, rowb, repeat=512):
    rng = np.random.default_rng(2)
    mode = SWZ_64B if rowb == 64 else SWZ_128B
    esz = 2 if kind == 0 else 4
    chans = rowb // esz
    rows = 256 + 64
    if kind == 0:
        X = (0.01 * rng.standard_normal((rows, chans))).astype(np.float16)
    else:
        X = tf32_round((0.01 * rng.standard_normal((rows, chans))).astype(np.float32))
    A_OFF, B_OFF = 0, ((rows * rowb + 1023) // 1024) * 1024
    img = Image(2 * B_OFF)
    img.put_rows(A_OFF, X.view(np.uint8).reshape(rows, rowb), mode)
    img.put_rows(B_OFF, X.view(np.uint8).reshape(rows, rowb), mode)
    if not mn_major:
        adesc = make_desc(A_OFF, 16, 8 * rowb, mode)
        bdesc = make_desc(B_OFF, 16, 8 * rowb, mode)
        n_k, step = rowb // 32, 32
    else:
        adesc = make_desc(A_OFF, rowb, 8 * rowb, mode)     # overlapping blocks one row apart
        bdesc = make_desc(B_OFF, rowb, 8 * rowb, mode)
        n_k, step = 4, 16 * rowb                          # advance 16 K rows per MMA
    idesc = make_idesc(FMT_F16 if kind == 0 else FMT_TF32, 128, N, int(mn_major), int(mn_major))
    _, c1, st1 = run_mma(img, adesc, bdesc, idesc, kind, n_k, step, step, N, repeat=1)
    _, c2, st2 = run_mma(img, adesc, bdesc, idesc, kind, n_k, step, step, N, repeat=repeat)
    per = (c2 - c1) / float(n_k * (repeat - 1))
    kk = 16 if kind == 0 else 8
    record("H3_cycles", kind=kind, N=N, mn_major=mn_major, rowb=rowb, cycles_per_mma=per,
           macs_per_cycle=128 * N * kk / per, status=st1 | st2)


# ------------------------------------------------------------------------------------------------ H4
def test_tma(elem, swizzle_enum, mode, coords, dst_off, box_w=12, box_h=20):
    dev = torch.device("cuda")
    D, H, W, C = 6, 24, 20, 32
    rng = np.random.default_rng(3)
    if elem == 2:
        x = rng.standard_normal((D, H, W, C)).astype(np.float16)
    else:
        x = rng.standard_normal((D, H, W, C)).astype(np.float32)
    xt = torch.from_numpy(x).to(dev)
    dims = np.array([C, W, H, D, 1], dtype=np.uint64)
    strides = np.array([C * elem, W * C * elem, H * W * C * elem, D * H * W * C * elem], dtype=np.uint64)
    box = np.array([C, box_w, box_h, 1, 1], dtype=np.uint32)
    crd = np.array(coords, dtype=np.int32)
    rowb = C * elem
    nbytes = box_w * box_h * rowb
    dump_bytes = ((dst_off + nbytes + 1023) // 1024) * 1024
    dump = torch.zeros(dump_bytes, dtype=torch.uint8, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    rc = lib.probe_tma(xt.data_ptr(), elem, dims.ctypes.data, strides.ctypes.data, box.ctypes.data, swizzle_enum,
                       crd.ctypes.data, dst_off, nbytes, dump.data_ptr(), dump_bytes, st.data_ptr())
    if rc != 0:
        record("H4_tma", elem=elem, swizzle=swizzle_enum, coords=list(map(int, coords)), dst_off=dst_off,
               error=lib.probe_last_error().decode(), ok=False)
        return
    got = dump.cpu().numpy()
    # expected logical box content with zero fill
    ref = np.zeros((box_h, box_w, C), dtype=x.dtype)
    _, w0, h0, d0, _ = coords
    for hh in range(box_h):
        for ww in range(box_w):
            h, w = h0 + hh, w0 + ww
            if 0 <= h < H and 0 <= w < W and 0 <= d0 < D:
                ref[hh, ww] = x[d0, h, w]
    rows = ref.view(np.uint8).reshape(box_h * box_w, rowb)
    out = {}
    for label, rel in (("abs", False), ("rel", True)):
        exp = np.full(dump_bytes, 0xEE, dtype=np.uint8)
        la = (np.arange(rows.shape[0])[:, None] * rowb + np.arange(rowb)[None, :])
        pa = (swz(la, mode) + dst_off) if rel else swz(la + dst_off, mode)
        exp[pa] = rows
        out[label] = bool(np.array_equal(exp, got))
    record("H4_tma", elem=elem, swizzle=swizzle_enum, coords=list(map(int, coords)), dst_off=dst_off,
           status=int(st.item()), match_abs_model=out["abs"], match_rel_model=out["rel"],
           ok=bool(out["abs"] or out["rel"]))


def main():
    assert torch.cuda.is_available(), "needs a GPU"
    print(torch.cuda.get_device_name(0))
    # H4 first (cheapest to interpret)
    for coords in ([0, -2, -2, 1, 0], [0, 10, 6, 5, 0], [0, 3, 1, 0, 0]):
        test_tma(2, 2, SWZ_64B, coords, 0)
        test_tma(4, 3, SWZ_128B, coords, 0)
    test_tma(2, 2, SWZ_64B, [0, 3, 1, 0, 0], 1024)
    test_tma(2, 2, SWZ_64B, [0, 3, 1, 0, 0], 256)
    test_tma(4, 3, SWZ_128B, [0, 3, 1, 0, 0], 512)
    test_tma(2, 2, SWZ_64B, [0, -2, -2, 1, 0], 0, box_w=16, box_h=20)
    # H1
    for kind, chans in ((0, 32), (0, 64), (1, 32)):
        for G in (8, 12, 16):
            for s in (0, 1, 2, 3, 5):
                for bo in (0, 1):
                    if s == 0 and bo == 1:
                        continue
                    try:
                        test_kmajor(kind, chans, s, G, bo)
                    except Exception as e:  # noqa: BLE001
                        record("H1_kmajor", kind=kind, chans=chans, shift=s, G=G, bo_mode=bo, error=str(e), ok=False)
    # H2
    for swap in (False, True):
        for (sa, sb, G) in ((8, 8, 8), (1, 1, 8), (12, 1, 12), (1, 1, 12), (16, 1, 16), (0, 1, 12)):
            try:
                test_mnmajor(sa, sb, G, swap)
            except Exception as e:  # noqa: BLE001
                record("H2_mnmajor", sh_a=sa, sh_b=sb, G=G, swap=swap, error=str(e), ok=False)
    test_mnmajor(12, 1, 12, False, NB=1)
    test_mnmajor(12, 1, 12, False, NB=4)
    # H3
    for kind, rowb in ((0, 64), (0, 128), (1, 128)):
        for N in (32, 64, 128, 160, 256):
            bench_shape(kind, N, False, rowb)
    for N in (32, 128, 160, 256):
        bench_shape(0, N, True, 64)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_umma.json", "w") as f:
        json.dump(RESULTS, f, indent=1)
    bad = [r for r in RESULTS if r.get("ok") is False]
    print(f"{len(RESULTS)} probes, {len(bad)} not ok")


if __name__ == "__main__":
    sys.exit(main())
