#!/usr/bin/env python
"""Headline benchmark: voxels/s of one MoDE-conv block (MoDEConv(5, 12, 32, 32), train mode) forward+backward
on a 1x32x32x128x128 synthetic volume per GPU -- BASELINE.json's metric ("voxels/sec MoDE-conv fwd+bwd
@32x128x128x32ch").  One JSON line on rank 0; contract in the task statement / DESIGN.md section "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = K1 re-param + conv forward (+BN statistics/apply/ReLU) + BN backward + dgrad + wgrad + K1b for one
volume per rank (weak scaling: volumes are independent; the only exchange is the data-parallel gradient
all-reduce of the layer's 0.6 MB of parameters over NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CI = CO = 32
D, H, W = 32, 128, 128
T = 12
VOX = D * H * W
FLOP_CONV = 2.0 * 125 * CI * CO * VOX            # one of fwd / dgrad / wgrad
METRIC = "voxels/sec MoDE-conv fwd+bwd @32x128x128x32ch"
TRAFFIC_FILES = ("r1_final_traffic.json", "r1e_traffic.json", "r1b_traffic.json")   # newest ncu --set full capture of the
TRAFFIC_FILE = next((f for f in TRAFFIC_FILES                                    # same command (tools/ncu_summarise.py)
                     if os.path.exists(os.path.join(ROOT, "profiles", f))), TRAFFIC_FILES[-1])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_fwd_bwd(steps, warmup):
    """The reference's CPU implementation of the path, as ported in oracle/mode_torch.py (the reference is a
    PyTorch program; /root/reference is not present on the GPU box). All host threads."""
    from oracle import mode_torch as otc
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    p = otc.init_mode_conv_params(T, CI, CO, generator=g)
    p = {k: (v.requires_grad_(True) if v.dtype.is_floating_point and "running" not in k and "pool" not in k else v)
         for k, v in p.items()}
    x = torch.randn(1, CI, D, H, W, generator=g).requires_grad_(True)
    dout = torch.randn(1, CO, D, H, W, generator=g)
    t = torch.tensor([3])
    times = []
    for i in range(warmup + steps):
        for v in p.values():
            if v.grad is not None:
                v.grad = None
        x.grad = None
        t0 = time.perf_counter()
        y = otc.mode_conv(p, "", x, t, True)
        y.backward(dout)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    times, cores = cpu_fwd_bwd(steps, max(1, min(args.warmup, 2)))
    tot = sum(times)
    val = VOX * len(times) / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MoDEConv(5,12,32,32) train fwd+bwd, x[1,32,32,128,128]", "device": "host CPU"},
        "cpu_baseline": {"value": val, "unit": "voxels/s", "cores": cores, "kind": "port",
                         "sample": f"{len(times)} full fwd+bwd steps of the same workload (oracle/mode_torch.py)"},
        "e2e": {"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().strip().splitlines():
            c = [v.strip() for v in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from repmode_b200 import functional as Fm, lib as L, parallel as par
    from repmode_b200.nn_modules import MoDEConv

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = L.load()
    torch.manual_seed(0)
    m = MoDEConv(5, T, CI, CO).to(dev).train()
    params = [p for p in m.parameters()]
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    x_host = torch.randn(1, CI, D, H, W, generator=g).pin_memory()                      # what a caller holds (NCDHW)
    x_dev = x_host.to(dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    dout = torch.randn(1, CO, D, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last_3d)
    task = torch.tensor([(3 * rank) % T], device=dev, dtype=torch.int32)
    stream = torch.cuda.current_stream()

    def eager_step(xin):
        for p in params:
            p.grad = None
        xin.grad = None
        y = m(xin, task)
        y.backward(dout)

    class GraphStep:
        """One training step of the block (forward + backward, all kernels of the path) captured ONCE into a CUDA graph on
        a static input buffer and replayed: the launch-bound Python/ctypes enqueue (~0.45 ms per step, as long as the GPU
        work itself) leaves the timed loop.  Gradients land in the static tensors the capture allocated."""

        def __init__(self, xin):
            self.x = xin
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):                                  # warm-up off the capture (allocator, lazy inits)
                    eager_step(xin)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for p in params:
                p.grad = None
            xin.grad = None
            self.graph = torch.cuda.CUDAGraph()
            l0 = lib.mode_launch_count()
            # with NCCL initialised its watchdog thread polls events while we capture: only this thread's calls are policed
            mode = {"capture_error_mode": "thread_local"} if world > 1 else {}
            with torch.cuda.graph(self.graph, **mode):
                y = m(xin, task)
                y.backward(dout)
            self.launches = lib.mode_launch_count() - l0        # kernels of this library inside one replay
            self.grads = [p.grad for p in params]
            self.bias_grad = m.gate.bias.grad

        def replay(self):
            self.graph.replay()
            for p, gr in zip(params, self.grads):
                p.grad = gr

    use_graph = os.environ.get("REPMODE_BENCH_GRAPH", "1") == "1"
    graph_note = ""

    def make_graph_step(xin):
        """GraphStep, or None (eager launches) when the capture is refused -- never a silent change of the work done."""
        nonlocal use_graph, graph_note
        if not use_graph:
            return None
        try:
            return GraphStep(xin)
        except Exception as e:  # noqa: BLE001
            use_graph = False
            graph_note = f" (CUDA-graph capture failed, eager launches: {type(e).__name__}: {str(e)[:120]})"
            torch.cuda.synchronize()
            for p in params:
                p.grad = None
            return None

    gs_res = make_graph_step(x_dev)

    def step_resident():
        if gs_res is not None:
            gs_res.replay()
        else:
            eager_step(x_dev)
        if world > 1:
            par.sync_gradients(params)            # one flat NCCL all-reduce of the block's 0.6 MB of gradients

    # e2e: the caller holds pinned NCDHW host tensors; every step's input crosses PCIe inside the timed region.
    # Like any input pipeline the copy of step i+1 is issued (side stream, double buffer) before step i computes;
    # the first copy of a timed run is not overlapped and nothing is copied for a step that is not run.
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty(x_host.shape, device=dev).requires_grad_(True), torch.empty(x_host.shape, device=dev).requires_grad_(True)]
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_used = [torch.cuda.Event(), torch.cuda.Event()]
    fast_level = int(os.environ.get("REPMODE_BENCH_FAST", "0"))  # profiling runs (ncu): 1 = skip the e2e and CPU legs,
    fast = fast_level >= 1                                       # 2 = also skip the per-kernel roofline timings
    gs_e2e = [make_graph_step(b) for b in xbuf] if (use_graph and not fast) else None
    if gs_e2e is not None and any(g_ is None for g_ in gs_e2e):
        gs_e2e = None

    def issue_copy(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_used[b])
            with torch.no_grad():
                xbuf[b].copy_(x_host, non_blocking=True)
            ev_copied[b].record(copy_stream)

    def run_e2e(steps):
        for b in (0, 1):
            ev_used[b].record(stream)
        issue_copy(0)
        for i in range(steps):
            b = i & 1
            if i + 1 < steps:
                issue_copy(i + 1)
            stream.wait_event(ev_copied[b])
            if gs_e2e is not None:
                gs_e2e[b].replay()
            else:
                eager_step(xbuf[b])
            ev_used[b].record(stream)
            if world > 1:
                par.sync_gradients(params)
            _ = m.gate.bias.grad.cpu()                                                    # D2H read of a step result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.mode_launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.mode_launch_count() - l0
        if gs_res is not None and fn is step_resident:
            launches = gs_res.launches * steps                  # replays do not pass through the host-side counter
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, launches

    warmup = max(3, args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches = timed(step_resident, args.steps, warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e = float("nan")
    if not fast:
        run_e2e(warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_e2e(args.steps)
        e1.record(stream)
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms_e2e = float(tms.item())

    if rank != 0:
        return
    if fast_level >= 2:                                          # launch-list runs: only the steps themselves
        print(json.dumps({"metric": METRIC, "value": world * VOX * args.steps / (ms * 1e-3), "unit": "voxels/s",
                          "ms_per_step": ms / args.steps, "profiling_run": True}), flush=True)
        return
    value = world * VOX * args.steps / (ms * 1e-3)
    e2e = world * VOX * args.steps / (ms_e2e * 1e-3) if not fast else None       # profiling runs skip the e2e leg

    # ---- per-kernel roofline of the dominant kernels, timed alone with CUDA events on the launch stream
    hbm_gbs, tf_burst, tf_sust, peak_kind = peaks()
    kern = {}
    use_umma = Fm.umma_shape_ok(CI, CO, D, H, W) and Fm.default_precision() == "f16"
    dtype = L.MODE_F16 if use_umma else L.MODE_F32
    with torch.no_grad():
        layer, ci, co = Fm._layer(*m._params())
        xn = Fm.to_ndhwc(x_dev.detach())
        dyn = Fm.to_ndhwc(dout)
        x_op = Fm.cast_f16(xn) if use_umma else xn
        dy_op = Fm.cast_f16(dyn) if use_umma else dyn
        su = torch.zeros(1, dtype=torch.int32, device=dev)
        gq, w_fwd, w_dg = Fm.reparam_fwd(layer, task, 1, ci, co, dtype, True, Fm.W_SCALE_F16 if use_umma else 1.0)

        def tk(fn, it=10):
            fn(); fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(it):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / it
        kern["conv_fwd_ms"] = tk(lambda: Fm.conv3d(x_op, dtype, w_fwd, su, 1, D, H, W, ci, co))
        kern["conv_dgrad_ms"] = tk(lambda: Fm.conv3d(dy_op, dtype, w_dg, su, 1, D, H, W, co, ci))
        if use_umma and not Fm.UMMA_WGRAD:
            kern["wgrad_ms"] = tk(lambda: Fm.conv3d_wgrad(xn, dyn, L.MODE_F32, 1, D, H, W, ci, co), it=3)
        else:
            kern["wgrad_ms"] = tk(lambda: Fm.conv3d_wgrad(x_op, dy_op, dtype, 1, D, H, W, ci, co))
        kern["reparam_fwd_ms"] = tk(lambda: Fm.reparam_fwd(layer, task, 1, ci, co, dtype, True, 256.0), it=50)
        # K1 on the widest layer of the U-Net (bottle_block.conv2, 512 -> 512: 163 MB of experts) -- the shape at
        # which the re-param kernel is actually HBM-bound (at 32 -> 32 it moves 1 MB and is launch/latency bound)
        big = MoDEConv(5, T, 512, 512).to(dev)
        lb, cb, ob = Fm._layer(*big._params())
        kern["reparam_fwd_512_ms"] = tk(lambda: Fm.reparam_fwd(lb, task, 1, cb, ob, dtype, False, 256.0), it=10)
        bytes_512 = 620.0 * 512 * 512 + 125.0 * 512 * 512 * (2 if use_umma else 4)
        del big
    conv_ms = kern["conv_fwd_ms"]
    achieved = FLOP_CONV / (conv_ms * 1e-3) / 1e12
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            tr = json.load(f)
        key = "conv3d_pair_kernel<1>" if use_umma else "conv3d_simt_kernel"
        traffic = sum(t["dram_bytes"] for t in tr[key]) / len(tr[key])
    except Exception:  # noqa: BLE001
        traffic = None
    roofline = {"kernel": "conv3d_pair_kernel<true> (K2 forward, tcgen05.mma.cta_group::2)" if use_umma else "conv3d_simt_kernel (K2 forward)",
                "bound": "tensor", "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s",
                "frac": achieved / tf_burst, "traffic": traffic,
                "traffic_source": "ncu --set full dram__bytes_read+write per launch, profiles/" + TRAFFIC_FILE,
                "peak_source": f"MEASURED_PEAKS.json bf16 burst ({peak_kind})",
                "algorithmic_flop_per_launch": FLOP_CONV,
                "others": {"dgrad_TFLOPs": FLOP_CONV / (kern["conv_dgrad_ms"] * 1e-3) / 1e12,
                           "wgrad_TFLOPs": FLOP_CONV / (kern["wgrad_ms"] * 1e-3) / 1e12,
                           "reparam_fwd_GBs": (620.0 * CI * CO + 2 * 125 * CI * CO * (2 if use_umma else 4)) /
                           (kern["reparam_fwd_ms"] * 1e-3) / 1e9,
                           "reparam_fwd_512x512_GBs": bytes_512 / (kern["reparam_fwd_512_ms"] * 1e-3) / 1e9,
                           "reparam_fwd_512x512_frac_of_hbm": bytes_512 / (kern["reparam_fwd_512_ms"] * 1e-3) / 1e9 / hbm_gbs,
                           "hbm_peak_GBs": hbm_gbs, **{k: round(v, 4) for k, v in kern.items()}}}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not fast:
        times, cores = cpu_fwd_bwd(3, 1)
        cpu = {"value": VOX * len(times) / sum(times), "unit": "voxels/s", "cores": cores, "kind": "port",
               "sample": "3 full fwd+bwd steps of the same workload after 1 warm-up (oracle/mode_torch.py, fp32)"}

    line = {
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands (power-of-two scaled, 10-bit mantissa like TF32), f32 accumulate" if use_umma else "f32",
        "data": "synthetic",
        "config": {"workload": "MoDEConv(5,12,32,32) train fwd+bwd, x[1,32,32,128,128] per GPU",
                   "layout": "NDHWC (channels_last_3d) resident", "parallelism": f"dp{world}",
                   "l2": "per-step working set ~0.5 GB > 126 MB L2 (inputs larger than L2, no explicit flush)",
                   "precision": Fm.default_precision(),
                   "launch": ("forward+backward captured once as a CUDA graph and replayed" if use_graph else "eager")
                   + graph_note},
        "e2e": {"value": e2e, "unit": "voxels/s", "h2d_bytes_per_step": x_host.numel() * 4 * world,
                "d2h_bytes_per_step": m.gate.bias.numel() * 4 * world,
                "ms_per_step": ms_e2e / args.steps if not fast else None,
                "api": "MoDEConv.forward(x from pinned NCDHW host memory, double-buffered H2D on a side stream) + "
                       "backward" + (" (CUDA-graph replay per input buffer)" if use_graph else "")
                       + ", gate.bias.grad.cpu() every step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MoDE-conv path has no CPU fallback); use --impl reference")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
