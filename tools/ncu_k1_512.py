"""The re-parameterisation kernels on the layer where they are bandwidth-bound (bottle_block.conv2, 512 -> 512: 163 MB of
experts) -- a handful of launches of each for `ncu --set full` (tools/gpu_r2k1.sh): K1 forward pack, K1 + dgrad pack, K1b +
gate backward.  Timings come from tools/bench_k1.py (CUDA events, no profiler); this exists for the DRAM byte counts and
the achieved-bandwidth percentages of the captures."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repmode_b200 import functional as Fm, lib as L  # noqa: E402
from repmode_b200.nn_modules import MoDEConv  # noqa: E402


def main():
    lib = L.load()
    ci = co = 512
    torch.manual_seed(0)
    m = MoDEConv(5, 12, ci, co).cuda()
    layer, _, _ = Fm._layer(*m._params())
    task = torch.tensor([3], device="cuda", dtype=torch.int32)
    su = torch.zeros(1, dtype=torch.int32, device="cuda")
    g = None
    for _ in range(2):
        g, w, _ = Fm.reparam_fwd(layer, task, 1, ci, co, L.MODE_F16, False, 256.0)       # K1 alone
    for _ in range(2):
        g, w, wd = Fm.reparam_fwd(layer, task, 1, ci, co, L.MODE_F16, True, 256.0)       # K1 + dgrad pack
    dweff = torch.randn(1, 125, co, ci, device="cuda")
    outs = [torch.empty_like(t) for t in m._params()]
    ws = torch.empty(max(int(lib.mode_reparam_bwd_workspace_bytes(ci, co, 1)), 16), dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        L.check(lib.mode_reparam_bwd(ctypes.byref(layer), Fm._p(task), None, 1, Fm._p(su), 1, Fm._p(g), Fm._p(dweff),
                                     *[Fm._p(o) for o in outs], Fm._p(ws), st), "mode_reparam_bwd")
    torch.cuda.synchronize()
    L.poll_error("ncu_k1_512")
    print("done")


if __name__ == "__main__":
    main()
