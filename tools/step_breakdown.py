"""Diagnostics: where one headline training step (MoDEConv(5,12,32,32) fwd+bwd on [1,32,32,128,128]) spends its time --
CPU enqueue time vs GPU time, per-kernel durations (torch.profiler / CUPTI, warm), and the same step replayed as a CUDA
graph.  python tools/step_breakdown.py -> gpurun_out/step_breakdown.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repmode_b200.nn_modules import MoDEConv  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = MoDEConv(5, 12, 32, 32).to(dev).train()
    params = list(m.parameters())
    x = torch.randn(1, 32, 32, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    dout = torch.randn(1, 32, 32, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
    task = torch.tensor([3], device=dev, dtype=torch.int32)

    def step():
        for p in params:
            p.grad = None
        x.grad = None
        y = m(x, task)
        y.backward(dout)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    out = {}
    n = 30
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    t_enq = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    out["cpu_enqueue_ms_per_step"] = 1e3 * t_enq / n
    out["gpu_ms_per_step"] = e0.elapsed_time(e1) / n
    # per-kernel (warm) durations
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        if ev.device_time_total > 0 and ev.device_type is not None and "cuda" in str(ev.device_type).lower():
            rows.append((ev.key[:70], ev.count / 5, ev.device_time_total / 5))
    rows.sort(key=lambda r: -r[2])
    out["kernels_us_per_step"] = [{"name": k, "launches": c, "us": round(t, 2)} for k, c, t in rows[:30]]
    out["kernel_sum_us_per_step"] = sum(r[2] for r in rows)
    # CUDA graph of the step
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        for p in params:
            p.grad = None
        x.grad = None
        with torch.cuda.graph(g):
            y = m(x, task)
            y.backward(dout)
        torch.cuda.synchronize()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out["graph_ms_per_step"] = e0.elapsed_time(e1) / n
        ref = [p.grad.clone() for p in params]
        step_ok = all(torch.isfinite(r).all().item() for r in ref)
        out["graph_grads_finite"] = bool(step_ok)
    except Exception as e:  # noqa: BLE001
        out["graph_error"] = repr(e)[:300]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "step_breakdown.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "kernels_us_per_step"}))
    for r in out["kernels_us_per_step"]:
        print(f"{r['us']:9.2f} us x{r['launches']:<5.1f} {r['name']}")


if __name__ == "__main__":
    main()
