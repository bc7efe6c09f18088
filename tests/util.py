"""Shared helpers for the parity tests: golden-fixture loading and the tolerance metrics
(max|d|/max|ref| and relative L2, SURVEY.md section 8c numerical notes)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    if "np_inputs" in d and bool(d["np_inputs"]):
        seed = int(d["seed"])
        shape = tuple(int(v) for v in d["x_shape"])
        d["x"] = np.random.RandomState(seed).standard_normal(shape).astype(np.float32)
        co = d["p.expert_conv5x5_conv"].shape[0]
        oshape = (shape[0], co) + shape[2:]
        d["dout"] = np.random.RandomState(seed + 1).standard_normal(oshape).astype(np.float32)
    return d


def params_of(d, prefix="p."):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def max_rel(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def rel_l2(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def assert_close(got, ref, tol, what=""):
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    m, l2 = max_rel(got, ref), rel_l2(got, ref)
    assert m <= tol and l2 <= tol, f"{what}: max_rel={m:.3e} rel_l2={l2:.3e} > tol={tol:.1e}"


def pack_weights(w_eff, half, dgrad=False):
    """numpy restatement of K1's packed layouts (include/repmode_b200.h): w_eff [U,Co,Ci,5,5,5] ->
    fp32 pack [U][tap][k_chunk][rows][32]; fp16 pack [U][k_chunk][kh*5+kw][4-kd][rows][32] with each block the
    64-byte-swizzled image (16-byte chunk index ^= (row >> 1) & 3). dgrad: flipped taps, rows=ci, k=co."""
    u, co, ci = w_eff.shape[:3]
    w = w_eff.reshape(u, co, ci, 125)
    if dgrad:
        w = w[..., ::-1].transpose(0, 2, 1, 3)          # [U, rows=ci, k=co, 124-tap]
    rows, k = w.shape[1], w.shape[2]
    nch = (k + 31) // 32
    out = np.zeros((u, 125, nch, rows, 32), dtype=np.float16 if half else np.float32)
    for c in range(nch):
        kk = min(32, k - c * 32)
        blk = w[:, :, c * 32:c * 32 + kk, :].transpose(0, 3, 1, 2)      # [U,125,rows,kk]
        if half:
            col = np.arange(kk)
            r = np.arange(rows)
            pc = ((((col[None, :] >> 3) ^ (r[:, None] >> 1)) & 3) << 3) | (col[None, :] & 7)   # [rows,kk]
            out[:, :, c, r[:, None], pc] = blk
        else:
            out[:, :, c, :, :kk] = blk
    if half:   # [U,125=(kd,t),nch,rows,32] -> [U,nch,t,4-kd,rows,32]
        out = out.reshape(u, 5, 25, nch, rows, 32)[:, ::-1].transpose(0, 3, 2, 1, 4, 5)
        out = np.ascontiguousarray(out)
    return out.reshape(-1)
