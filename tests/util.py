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
