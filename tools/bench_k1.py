"""K1 / K1b (re-parameterisation forward / backward) against the HBM roofline, per layer shape of the U-Net.
Algorithmic bytes: fwd = 620*Co*Ci + U*125*Co*Ci*s (s = 2 fp16; x2 with the dgrad pack);
bwd = 500*N*Co*Ci (d_weff) + 2*620*Co*Ci.   python tools/bench_k1.py -> gpurun_out/bench_k1.json"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repmode_b200 import functional as Fm, lib as L  # noqa: E402
from repmode_b200.nn_modules import MoDEConv  # noqa: E402


def tk(fn, it=20):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def main():
    lib = L.load()
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        hbm = 6650.0
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = ((32, 32), (64, 64), (128, 128), (256, 256), (512, 256), (512, 512))
    if os.environ.get("K1_ONLY"):
        c = int(os.environ["K1_ONLY"])
        shapes = ((c, c),)
    for ci, co in shapes:
        torch.manual_seed(0)
        m = MoDEConv(5, 12, ci, co).cuda()
        layer, _, _ = Fm._layer(*m._params())
        task = torch.tensor([3], device="cuda", dtype=torch.int32)
        su = torch.zeros(1, dtype=torch.int32, device="cuda")
        g, w, wd = Fm.reparam_fwd(layer, task, 1, ci, co, L.MODE_F16, True, 256.0)
        dweff = torch.randn(1, 125, co, ci, device="cuda")
        outs = [torch.empty_like(t) for t in m._params()]
        ws = torch.empty(max(int(lib.mode_reparam_bwd_workspace_bytes(ci, co, 1)), 16), dtype=torch.uint8, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def bwd():
            L.check(lib.mode_reparam_bwd(ctypes.byref(layer), Fm._p(task), None, 1, Fm._p(su), 1, Fm._p(g), Fm._p(dweff),
                                         *[Fm._p(o) for o in outs], Fm._p(ws), st), "mode_reparam_bwd")
        # L2 flush between launches (the experts of the small layers would otherwise sit in the 126 MB L2)
        t_flush = tk(lambda: flush.zero_())
        t_f = tk(lambda: (flush.zero_(), Fm.reparam_fwd(layer, task, 1, ci, co, L.MODE_F16, False, 256.0))) - t_flush
        t_fd = tk(lambda: (flush.zero_(), Fm.reparam_fwd(layer, task, 1, ci, co, L.MODE_F16, True, 256.0))) - t_flush
        t_b = tk(lambda: (flush.zero_(), bwd())) - t_flush
        bf = 620.0 * co * ci + 125.0 * co * ci * 2
        bfd = bf + 125.0 * co * ci * 2 * 2          # the pack pass re-reads the fwd pack and writes the dgrad pack
        bb = 500.0 * co * ci + 2 * 620.0 * co * ci
        out[f"{ci}x{co}"] = {"fwd_us": t_f * 1e3, "fwd_GBs": bf / t_f / 1e6, "fwd_frac_hbm": bf / t_f / 1e6 / hbm,
                             "fwd_dgrad_us": t_fd * 1e3, "fwd_dgrad_GBs": bfd / t_fd / 1e6,
                             "bwd_us": t_b * 1e3, "bwd_GBs": bb / t_b / 1e6, "bwd_frac_hbm": bb / t_b / 1e6 / hbm}
        print(f"{ci:4d}->{co:4d}  K1 fwd {t_f * 1e3:7.1f} us {bf / t_f / 1e6:7.0f} GB/s ({bf / t_f / 1e6 / hbm:5.1%})   "
              f"fwd+dgrad pack {t_fd * 1e3:7.1f} us {bfd / t_fd / 1e6:7.0f} GB/s   K1b {t_b * 1e3:7.1f} us "
              f"{bb / t_b / 1e6:7.0f} GB/s ({bb / t_b / 1e6 / hbm:5.1%})", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_k1.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
