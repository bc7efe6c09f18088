"""Diagnostics: where the whole-U-Net forward (eval, BASELINE.json config 2) and train step (config 3) spend their GPU
time -- per MoDEConv call site (CUDA events around every MoDEConv.forward / backward is not separable, so backward is
reported per kernel) and per kernel (torch.profiler).   python tools/profile_net.py [--train] [--batch 1]"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    a = ap.parse_args()
    os.environ["REPMODE_EVAL_GRAPH"] = "0"
    from repmode_b200 import nn_modules as NM
    mod = importlib.import_module("fnet.nn_modules.RepMode")
    torch.manual_seed(0)
    net = mod.Net(argparse.Namespace(adopted_datasets=list(range(12)), gpu_ids=0)).cuda()
    B = a.batch
    x = torch.randn(B, 1, 32, 128, 128, device="cuda")
    t = (torch.arange(B, device="cuda") * 3) % 12
    names = {m: n for n, m in net.named_modules() if isinstance(m, NM.MoDEConv)}
    events = []
    orig = NM.MoDEConv.forward

    def timed_forward(self, xx, tt, x2=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(self, xx, tt, x2)
        e1.record()
        shape = tuple(xx.shape) if x2 is None else (xx.shape[0], xx.shape[1] + x2.shape[1]) + tuple(xx.shape[2:])
        events.append((names[self], shape, self.out_chan, e0, e1))
        return out

    if a.train:
        net.train()
        tgt = torch.randn_like(x)

        def step():
            for p in net.parameters():
                p.grad = None
            loss = torch.mean((net(x, t) - tgt) ** 2)
            loss.backward()
    else:
        net.eval()

        def step():
            with torch.no_grad():
                net(x, t)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    NM.MoDEConv.forward = timed_forward
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    NM.MoDEConv.forward = orig
    print(f"one {'train' if a.train else 'eval'} step, batch {B}: {e0.elapsed_time(e1):.3f} ms (eager, with event overhead)")
    tot = 0.0
    for name, shape, co, a0, a1 in events:
        ms = a0.elapsed_time(a1)
        tot += ms
        ci = shape[1]
        gf = 2.0 * 125 * ci * co * shape[0] * shape[2] * shape[3] * shape[4] / 1e9
        print(f"  {name:38s} {ci:4d}->{co:4d} @{shape[2]:3d}x{shape[3]:3d}x{shape[4]:3d}  fwd {ms * 1e3:8.1f} us  {gf / ms:8.1f} TFLOP/s")
    print(f"  sum of MoDEConv forwards: {tot:.3f} ms")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in rows)
    print(f"kernel time per step: {total / 3 / 1e3:.3f} ms")
    for e in rows[:30]:
        print(f"  {e.device_time_total / 3:9.1f} us x{e.count / 3:5.1f}  {e.key[:100]}")


if __name__ == "__main__":
    main()
