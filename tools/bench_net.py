"""Whole-network timings (BASELINE.json configs[1] and configs[2]) through the plugin API; not the headline
bench.  python tools/bench_net.py [--batch 4] [--steps 5]  -> gpurun_out/bench_net.json"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--skip-train", action="store_true")
    a = ap.parse_args()
    import importlib
    from repmode_b200 import lib as L
    mod = importlib.import_module("fnet.nn_modules.RepMode")
    torch.manual_seed(0)
    net = mod.Net(argparse.Namespace(adopted_datasets=list(range(12)), gpu_ids=0)).cuda()
    out = {}
    vox = 32 * 128 * 128

    def timeit(fn, steps):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.load().mode_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, (L.load().mode_launch_count() - l0) / steps

    # configs[1]: eval forward, 1 x 32 x 128 x 128
    net.eval()
    x1 = torch.randn(1, 1, 32, 128, 128, device="cuda")
    t1 = torch.tensor([3], device="cuda")
    with torch.no_grad():
        ms, launches = timeit(lambda: net(x1, t1), a.steps)
    out["net_eval_fwd_b1"] = {"ms": ms, "voxels_per_s": vox / ms * 1e3, "tflops": 1086.3e9 / ms / 1e9, "our_launches": launches}
    print(json.dumps(out["net_eval_fwd_b1"]), flush=True)
    if not a.skip_train:
        # configs[2]: train step (fwd + bwd + Adam), batch B, distinct tasks
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        B = a.batch
        x = torch.randn(B, 1, 32, 128, 128, device="cuda")
        tgt = torch.randn(B, 1, 32, 128, 128, device="cuda")
        t = (torch.arange(B, device="cuda") * 3) % 12

        def step():
            opt.zero_grad(set_to_none=True)
            loss = torch.mean((net(x, t) - tgt) ** 2)
            loss.backward()
            opt.step()
        ms, launches = timeit(step, a.steps)
        out[f"net_train_step_b{B}"] = {"ms": ms, "voxels_per_s": B * vox / ms * 1e3,
                                       "tflops": 3 * B * 1086.3e9 / ms / 1e9, "our_launches": launches,
                                       "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        print(json.dumps(out[f"net_train_step_b{B}"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_net.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
