"""Diagnostics: where the tcgen05 conv kernel's MMA-issuing warp spends its cycles (mode_debug_profile), on the
headline 32->32 layer shape.  python tools/profile_conv.py  -> gpurun_out/conv_waits.json"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repmode_b200 import functional as Fm, lib as L  # noqa: E402
from tests.util import pack_weights  # noqa: E402
import numpy as np  # noqa: E402


def main():
    lib = L.load()
    out = []
    impls = [int(a) for a in sys.argv[1:]] or [3, 4]
    for impl, (n, d, h, w, k, nout) in [(i, sh) for sh in [(1, 32, 128, 128, 32, 32), (4, 32, 128, 128, 32, 32)] for i in impls]:
        x = torch.randn(n, d, h, w, k, device="cuda").half()
        weff = (np.random.RandomState(0).randn(n, nout, k, 5, 5, 5) * 0.02).astype(np.float32)
        w16 = torch.from_numpy(pack_weights(weff, half=True)).cuda()
        su = torch.arange(n, dtype=torch.int32, device="cuda")
        prof = torch.zeros(8 * 160, dtype=torch.int64, device="cuda")
        for _ in range(3):
            Fm.conv3d(x, L.MODE_F16, w16, su, n, d, h, w, k, nout, impl=impl)
        lib.mode_debug_profile(ctypes.c_void_p(prof.data_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.mode_debug_profile(None)
        e2.record()
        for _ in range(5):
            Fm.conv3d(x, L.MODE_F16, w16, su, n, d, h, w, k, nout, impl=impl)
        e3.record()
        torch.cuda.synchronize()
        ms_noprof = e2.elapsed_time(e3) / 5
        lib.mode_debug_profile(ctypes.c_void_p(prof.data_ptr()))
        e0.record()
        Fm.conv3d(x, L.MODE_F16, w16, su, n, d, h, w, k, nout, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        lib.mode_debug_profile(None)
        code = ctypes.c_int32(0)
        lib.mode_poll_error(ctypes.byref(code))
        ncta = 148 if impl != 4 else 74
        p = prof.view(160, 8)[:ncta].cpu().double()
        t_entry, t_mma, t_done = p[:, 4], p[:, 5], p[:, 6]
        t0g = float(t_entry.min())
        flop = 2.0 * 125 * k * nout * n * d * h * w
        r = {"impl": impl, "err": code.value, "shape": [n, d, h, w, k, nout], "ms_noprof": ms_noprof, "ms": e0.elapsed_time(e1), "tflops": flop / e0.elapsed_time(e1) / 1e9,
             "mma_warp_cycles_mean": float(p[:, 0].mean()), "mma_warp_cycles_max": float(p[:, 0].max()),
             "wait_tmem_mean": float(p[:, 1].mean()), "wait_weights_mean": float(p[:, 2].mean()),
             "issue_mean": float(p[:, 7].mean()), "wait_planes_mean": float(p[:, 3].mean()), "wait_planes_max": float(p[:, 3].max()),
             "entry_spread_us": float((t_entry.max() - t0g) / 1e3), "mma_end_us_mean": float((t_mma.mean() - t0g) / 1e3),
             "mma_end_us_max": float((t_mma.max() - t0g) / 1e3), "done_us_max": float((t_done.max() - t0g) / 1e3),
             "tail_after_mma_us_mean": float(((t_done - t_mma).mean()) / 1e3)}
        r["per_cta_cycles"] = [int(v) for v in p[:, 0].tolist()]
        r["per_cta_tmem_wait"] = [int(v) for v in p[:, 1].tolist()]
        r["per_cta_plane_wait"] = [int(v) for v in p[:, 3].tolist()]
        out.append(r)
        print(json.dumps({k: v for k, v in r.items() if not k.startswith('per_cta')}), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "conv_waits.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
