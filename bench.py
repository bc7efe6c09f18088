#!/usr/bin/env python
"""Headline benchmark: voxels/s of one MoDE-conv block (MoDEConv(5, 12, 32, 32), train mode) forward+backward on a
32x128x128 slab of 32 channels per GPU -- BASELINE.json's metric ("voxels/sec MoDE-conv fwd+bwd @32x128x128x32ch").
One JSON line on rank 0; contract in the task statement / DESIGN.md section "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--comm peer|nccl]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = K1 re-param + conv forward (+BN statistics / apply / ReLU) + BN backward + dgrad + wgrad + K1b.
  N = 1 : one volume x[1,32,32,128,128].
  N > 1 : the north-star split -- ONE volume x[1,32,32*N,128,128] sharded on the D axis, a 32-plane slab per GPU, so the
          per-GPU work is the headline shape (weak scaling).  Inside the timed step: halo exchange of the conv operand
          (2 planes per side) and of dy, BatchNorm statistics all-reduced over the owned planes (forward and backward), and
          the parameter gradients summed over the slabs (repmode_b200/sharded.py).  The exchange steps
          are stores into the neighbour's memory over NVLink (--comm peer, default) or NCCL (--comm nccl; also measured as
          the secondary field `nccl_value`); `replicas_value` is round 1's number (independent volumes + gradient all-reduce).
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
TRAFFIC_FILES = ("r2f_traffic.json", "r2_traffic.json", "r1_final_traffic.json", "r1e_traffic.json", "r1b_traffic.json")   # newest ncu --set full
TRAFFIC_FILE = next((f for f in TRAFFIC_FILES                                                          # capture first
                     if os.path.exists(os.path.join(ROOT, "profiles", f))), TRAFFIC_FILES[-1])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--comm", default=os.environ.get("REPMODE_BENCH_COMM", "peer"), choices=["peer", "nccl"])
    # extra lines for BASELINE.json configs 2-5 (the driver runs the default = headline): whole-U-Net workloads
    ap.add_argument("--config", default="headline", choices=["headline", "net_fwd", "net_train", "cfg4", "cfg5"])
    return ap.parse_args()


def workload_config(world):
    """The `config` object of the JSON line -- the same for our arm and the reference arm."""
    if world == 1:
        wl = "MoDEConv(5,12,32,32) train fwd+bwd, x[1,32,32,128,128] per GPU"
        par = "dp1"
    else:
        wl = (f"MoDEConv(5,12,32,32) train fwd+bwd, ONE volume x[1,32,{D * world},128,128] sharded on D: "
              "a [1,32,32,128,128] slab per GPU")
        par = f"d-shard x{world}: halo exchange (x, dy) + BatchNorm all-reduce (fwd, bwd) + gradient all-reduce per step"
    return {"workload": wl, "layout": "NDHWC (channels_last_3d) resident", "parallelism": par,
            "l2": "per-step working set ~0.5 GB > 126 MB L2 (inputs larger than L2, no explicit flush)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_fwd_bwd(steps, warmup, planes=D):
    """The reference's CPU implementation of the path, as ported in oracle/mode_torch.py (the reference is a
    PyTorch program; /root/reference is not present on the GPU box). All host threads."""
    from oracle import mode_torch as otc
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    p = otc.init_mode_conv_params(T, CI, CO, generator=g)
    p = {k: (v.requires_grad_(True) if v.dtype.is_floating_point and "running" not in k and "pool" not in k else v)
         for k, v in p.items()}
    x = torch.randn(1, CI, planes, H, W, generator=g).requires_grad_(True)
    dout = torch.randn(1, CO, planes, H, W, generator=g)
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded sample: the whole workload (all `world` slabs = the whole volume) when it stays within a few minutes
    # (~0.25 s per slab and step on 16 cores), else one slab per step, extrapolated per voxel
    planes = D * world if world * (steps + warmup) <= 240 else D
    times, cores = cpu_fwd_bwd(steps, warmup, planes)
    tot = sum(times)
    val = planes * H * W * len(times) / tot
    sample = (f"{len(times)} full fwd+bwd steps of the same workload (oracle/mode_torch.py, fp32)" if planes == D * world else
              f"{len(times)} fwd+bwd steps on ONE of the {world} slabs (oracle/mode_torch.py, fp32); voxels/s is per-voxel "
              "extrapolation-free (a CPU does not shard)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world), "reference_device": "host CPU",
        "cpu_baseline": {"value": val, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=100):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", str(period_ms), "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().strip().splitlines():
            c = [v.strip() for v in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from repmode_b200 import functional as Fm, lib as L, parallel as par, peer, sharded
    from repmode_b200.nn_modules import MoDEConv

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = L.load()
    torch.manual_seed(0)
    m = MoDEConv(5, T, CI, CO).to(dev).train()
    params = [p for p in m.parameters()]
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    x_host = torch.randn(1, CI, D, H, W, generator=g).pin_memory()                      # what a caller holds (NCDHW slab)
    x_dev = x_host.to(dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    dout = torch.randn(1, CO, D, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last_3d)
    # one volume -> one task (N > 1: every slab belongs to the same sample); N = 1 keeps round 1's task id
    task = torch.tensor([3 if world > 1 else 0], device=dev, dtype=torch.int32)
    stream = torch.cuda.current_stream()
    fast_level = int(os.environ.get("REPMODE_BENCH_FAST", "0"))  # profiling runs (ncu): 1 = skip the e2e / CPU / secondary
    fast = fast_level >= 1                                       # arms, 2 = also skip the per-kernel roofline timings

    comms = {}
    if world > 1:
        arena = 3 * (D + 4) * H * W * CI * 2 + (32 << 20)                               # haloed fp16 x, dy (+ a probe buffer) + slots
        comms["peer"] = peer.PeerComm(dev, arena) if (args.comm == "peer" or not fast) else None
        comms["nccl"] = peer.TorchComm() if (args.comm == "nccl" or not fast) else None

    def zero_grads(xin):
        for p in params:
            p.grad = None
        xin.grad = None

    def step_single(xin):
        zero_grads(xin)
        y = m(xin, task)
        y.backward(dout)

    def make_step_sharded(comm, tag):
        def step(xin):
            zero_grads(xin)
            y = sharded.sharded_mode_conv(m, xin, task, comm, D * world, tag=tag)
            y.backward(dout)
        return step

    class GraphStep:
        """One training step of the block (forward + backward, all kernels of the path, the peer-memory exchange steps
        included) captured ONCE into a CUDA graph on a static input buffer and replayed: the launch-bound Python/ctypes
        enqueue (~0.7 ms per step, more than the GPU work) leaves the timed loop.  Gradients land in the static tensors
        the capture allocated."""

        def __init__(self, fn, xin):
            self.x = xin
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):                                  # warm-up off the capture (allocator, lazy inits)
                    fn(xin)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            zero_grads(xin)
            self.graph = torch.cuda.CUDAGraph()
            l0 = lib.mode_launch_count()
            # with NCCL initialised its watchdog thread polls events while we capture: only this thread's calls are policed
            mode = {"capture_error_mode": "thread_local"} if world > 1 else {}
            with torch.cuda.graph(self.graph, **mode):
                fn(xin)
            self.launches = lib.mode_launch_count() - l0        # kernels of this library inside one replay
            self.grads = [p.grad for p in params]

        def replay(self):
            self.graph.replay()
            for p, gr in zip(params, self.grads):
                p.grad = gr

    use_graph = os.environ.get("REPMODE_BENCH_GRAPH", "1") == "1"
    notes = []

    def make_runner(fn, xin, label):
        """(callable running one step on `xin`, launches per step or None): a CUDA-graph replay, or eager launches when
        the capture is refused -- never a silent change of the work done."""
        if use_graph:
            try:
                gs = GraphStep(fn, xin)
                return gs.replay, gs.launches, True
            except Exception as e:  # noqa: BLE001
                notes.append(f"{label}: CUDA-graph capture failed, eager launches ({type(e).__name__}: {str(e)[:100]})")
                torch.cuda.synchronize()
                zero_grads(xin)
        return (lambda: fn(xin)), None, False

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
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, launches

    # ---- the headline arm: N = 1 single volume; N > 1 D-sharded volume through the selected comm
    if world == 1 and os.environ.get("REPMODE_BENCH_SHARDED_LOCAL", "0") == "1":
        # diagnostics: the D-sharded formulation (haloed operand buffers, conv on owned planes) with NO neighbour -- what the
        # slab geometry alone costs against the plain single-volume step
        main_fn, main_label = make_step_sharded(peer.TorchComm(), "blk"), "d-sharded formulation, no neighbours"
    elif world == 1:
        main_fn, main_label = step_single, "single volume"
    else:
        main_fn, main_label = make_step_sharded(comms[args.comm], "blk"), f"d-sharded, {args.comm}"
    run_main, graph_launches, main_graphed = make_runner(main_fn, x_dev, main_label)

    if os.environ.get("REPMODE_BENCH_PROFILE", "0") == "1":
        # diagnostics: per-kernel GPU time of the (eager) step on every rank, rank 0 prints; the peer wait kernels' durations
        # are the exposed exchange latency
        from torch.profiler import ProfilerActivity, profile
        for _ in range(3):
            main_fn(x_dev)
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                main_fn(x_dev)
            torch.cuda.synchronize()
        barrier()
        if rank == 0:
            rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
            for e in rows[:28]:
                print(f"{e.device_time_total / 5:9.2f} us x{e.count / 5:4.1f}  {e.key[:90]}", file=sys.stderr)

    warmup = max(3, args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches = timed(run_main, args.steps, warmup)
    clocks = sampler.stop() if sampler else None
    if graph_launches is not None:
        launches = graph_launches * args.steps                      # replays do not pass through the host-side counter
    value = world * VOX * args.steps / (ms * 1e-3)
    if fast_level >= 2:                                             # launch-list runs: only the steps themselves
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "voxels/s", "ms_per_step": ms / args.steps,
                              "profiling_run": True}), flush=True)
        return

    # ---- sustained sub-measurement: the same step back to back for >= 2 s (clocks settle below the burst clock)
    sustained = None
    if not fast:
        n_sus = int(min(20000, max(args.steps, 2.2 / max(1e-6, ms / args.steps * 1e-3))))
        s2 = ClockSampler(local_rank, 50) if rank == 0 else None
        ms_s, _ = timed(run_main, n_sus, 0)
        c2 = s2.stop() if s2 else None
        sustained = {"value": world * VOX * n_sus / (ms_s * 1e-3), "unit": "voxels/s", "steps": n_sus,
                     "seconds": ms_s * 1e-3, "ms_per_step": ms_s / n_sus, "clocks": c2}

    # ---- secondary arms at N > 1: the other comm back end, and round 1's independent replicas
    secondary = {}
    if world > 1 and not fast:
        other = "nccl" if args.comm == "peer" else "peer"
        run_o, _, graphed_o = make_runner(make_step_sharded(comms[other], "blk_" + other), x_dev, f"d-sharded, {other}")
        ms_o, _ = timed(run_o, args.steps, warmup)
        secondary[other + "_value"] = world * VOX * args.steps / (ms_o * 1e-3)
        secondary[other + "_ms_per_step"] = ms_o / args.steps
        secondary[other + "_launch"] = "cuda graph" if graphed_o else "eager"
        run_r, _, _ = make_runner(step_single, x_dev, "replicas")

        def step_replicas():
            run_r()
            par.sync_gradients(params)            # one flat NCCL all-reduce of the block's 0.6 MB of gradients
        ms_r, _ = timed(step_replicas, args.steps, warmup)
        secondary["replicas_value"] = world * VOX * args.steps / (ms_r * 1e-3)
        secondary["replicas_ms_per_step"] = ms_r / args.steps
        # per exchange step, timed alone on this rank's stream (microseconds, max over ranks)
        per = {}
        for kind, comm in comms.items():
            xe = comm.alloc(("bench", "x_ext"), (1, D + 4, H, W, CI), torch.float16, dev)
            v64 = torch.zeros(2 * CO, dtype=torch.float64, device=dev)
            gfl = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
            for name, fn in (("halo_2planes_fp16", lambda: comm.halo_fill(xe, 2, "bench.h")),
                             ("bn_allreduce_64xf64", lambda: comm.all_reduce(v64, "bench.bn")),
                             ("grad_allreduce_%dxf32" % gfl.numel(), lambda: comm.all_reduce(gfl, "bench.g"))):
                ms_c, _ = timed(fn, 20, 3)
                per[f"{kind}.{name}_us"] = round(1e3 * ms_c / 20, 2)
        secondary["exchange_step_us"] = per

    # ---- e2e: the caller holds pinned NCDHW host tensors; every step's input slab crosses PCIe inside the timed region
    # (copy of step i+1 issued on a side stream before step i computes, double buffer) and the step's RESULT -- all
    # parameter gradients of the block -- is read back to the host every step.
    ms_e2e = float("nan")
    d2h_bytes = sum(p.numel() for p in params) * 4
    if not fast:
        copy_stream = torch.cuda.Stream(device=dev)
        xbuf = [torch.empty(x_host.shape, device=dev).requires_grad_(True) for _ in range(2)]
        runners = [make_runner(main_fn, b, "e2e")[0] for b in xbuf]
        ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
        ev_used = [torch.cuda.Event(), torch.cuda.Event()]
        g_host = torch.empty(d2h_bytes // 4, dtype=torch.float32).pin_memory()

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
                runners[b]()
                ev_used[b].record(stream)
                off = 0
                for p in params:                                                          # D2H read of the step's result
                    n_ = p.numel()
                    g_host[off:off + n_].copy_(p.grad.reshape(-1), non_blocking=True)
                    off += n_
                stream.synchronize()

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
    traffic, traffic_wgrad = None, None
    try:
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            tr = json.load(f)
        key = "conv3d_pair_kernel<1>" if use_umma else "conv3d_simt_kernel"
        traffic = sum(t["dram_bytes"] for t in tr[key]) / len(tr[key])
        for kw in ("wgrad_deep_kernel", "wgrad_split_kernel"):
            if kw in tr:
                traffic_wgrad = sum(t["dram_bytes"] for t in tr[kw]) / len(tr[kw])
                break
    except Exception:  # noqa: BLE001
        pass
    wgrad_tf = FLOP_CONV / (kern["wgrad_ms"] * 1e-3) / 1e12
    step_flop = 3 * FLOP_CONV
    roofline = {"kernel": "conv3d_pair_kernel<true> (K2 forward, tcgen05.mma.cta_group::2)" if use_umma else "conv3d_simt_kernel (K2 forward)",
                "bound": "tensor", "achieved": achieved, "peak": tf_burst, "unit": "TFLOP/s",
                "frac": achieved / tf_burst, "traffic": traffic,
                "traffic_source": "ncu --set full dram__bytes_read+write per launch, profiles/" + TRAFFIC_FILE,
                "peak_source": f"MEASURED_PEAKS.json bf16 burst ({peak_kind}); kernels timed alone",
                "algorithmic_flop_per_launch": FLOP_CONV,
                "slowest_kernel": {"kernel": "wgrad_deep_kernel + reduce (K4)", "bound": "tensor", "achieved": wgrad_tf,
                                   "peak": tf_burst, "unit": "TFLOP/s", "frac": wgrad_tf / tf_burst,
                                   "traffic": traffic_wgrad},
                "whole_step": {"achieved": step_flop / (ms / args.steps * 1e-3) / 1e12 * 1.0, "peak": tf_burst, "unit": "TFLOP/s",
                               "frac": step_flop / (ms / args.steps * 1e-3) / 1e12 / tf_burst,
                               "frac_sustained": (step_flop / (sustained["ms_per_step"] * 1e-3) / 1e12 / tf_sust) if sustained else None,
                               "per_gpu": True},
                "others": {"dgrad_TFLOPs": FLOP_CONV / (kern["conv_dgrad_ms"] * 1e-3) / 1e12,
                           "wgrad_TFLOPs": wgrad_tf,
                           "reparam_fwd_GBs": (620.0 * CI * CO + 2 * 125 * CI * CO * (2 if use_umma else 4)) /
                           (kern["reparam_fwd_ms"] * 1e-3) / 1e9,
                           "reparam_fwd_512x512_GBs": bytes_512 / (kern["reparam_fwd_512_ms"] * 1e-3) / 1e9,
                           "reparam_fwd_512x512_frac_of_hbm": bytes_512 / (kern["reparam_fwd_512_ms"] * 1e-3) / 1e9 / hbm_gbs,
                           "hbm_peak_GBs": hbm_gbs, "bf16_sustained_TFLOPs": tf_sust,
                           **{k: round(v, 4) for k, v in kern.items()}}}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not fast:
        times, cores = cpu_fwd_bwd(3, 1)
        cpu = {"value": VOX * len(times) / sum(times), "unit": "voxels/s", "cores": cores, "kind": "port",
               "sample": "3 full fwd+bwd steps of the same workload after 1 warm-up (oracle/mode_torch.py, fp32)"}

    cfg = workload_config(world)                     # identical to the reference arm's `config`
    run_info = {"precision": Fm.default_precision(),
                "launch": "forward+backward captured once as a CUDA graph and replayed" if main_graphed else "eager"}
    if world > 1:
        run_info["comm"] = ("peer memory over NVLink (repmode_b200/csrc/peer.cu)" if args.comm == "peer"
                            else "NCCL (torch.distributed)")
    if notes:
        run_info["notes"] = notes
    line = {
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands (saturating, power-of-two scaled weights / dy; 10-bit mantissa like TF32), f32 accumulate" if use_umma else "f32",
        "data": "synthetic",
        "config": cfg, "run": run_info,
        "e2e": {"value": e2e, "unit": "voxels/s", "h2d_bytes_per_step": x_host.numel() * 4 * world,
                "d2h_bytes_per_step": d2h_bytes * world,
                "ms_per_step": ms_e2e / args.steps if not fast else None,
                "api": ("MoDEConv.forward" if world == 1 else "sharded_mode_conv") +
                       "(x slab from pinned NCDHW host memory, double-buffered H2D on a side stream) + backward, then every "
                       "parameter gradient of the block copied to pinned host memory and synchronised, each step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "sustained": sustained,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if world > 1:
        line["collectives_per_step"] = 5
        line.update(secondary)
    elif not fast and os.environ.get("REPMODE_BENCH_CONFIGS", "1") == "1":
        line["other_configs"] = other_configs_summary()
    print(json.dumps(line), flush=True)


def other_configs_summary():
    """BASELINE.json configs 2 and 3 (whole U-Net eval forward / train step) measured in the same run as the headline
    line, each by `bench.py --config ...` in a CHILD process under a timeout, so that nothing it does can cost the
    headline line: a compact {ms_per_step, value, e2e, conv_frac_of_peak} per config, or the reason it is missing."""
    import subprocess
    out = {}
    for name, steps in (("net_fwd", 20), ("net_train", 10)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", name, "--steps", str(steps),
                                "--warmup", "3"], capture_output=True, text=True, timeout=240)
            rec = None
            for ln in r.stdout.splitlines():
                if ln.startswith("{"):
                    rec = json.loads(ln)
            if rec is None:
                out[name] = {"unavailable": f"exit {r.returncode}: " + (r.stderr.strip().splitlines() or [""])[-1][:160]}
                continue
            out[name] = {"workload": rec["config"]["workload"], "ms_per_step": rec["ms_per_step"], "value": rec["value"],
                         "unit": rec["unit"], "e2e_value": rec["e2e"]["value"], "e2e_ms_per_step": rec["e2e"]["ms_per_step"],
                         "gpu_launches": rec.get("gpu_launches"), "conv_TFLOPs": rec["roofline"]["achieved"],
                         "conv_frac_of_bf16_burst": rec["roofline"]["frac"]}
        except Exception as e:  # noqa: BLE001 -- never lose the headline line to a secondary measurement
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return out


# ------------------------------------------------------------------------------------------- whole-U-Net configs
NET_FLOP_FWD = 1086.3e9        # MoDE convs of the full U-Net per 32x128x128 sample; fwd+bwd = 3x


def run_net_config(args, rank, local_rank, world):
    """BASELINE.json configs 2-5 through the plugin API (fnet.nn_modules.RepMode.Net):
       net_fwd   (config 2) eval forward, x[1,1,32,128,128], 1 GPU
       net_train (config 3) train step (fwd + bwd + Adam), batch 4 x 32x128x128, 4 distinct tasks, 1 GPU
       cfg4 / cfg5 (configs 4 / 5) train step on ONE volume 64x256x256 / 128x512x512, D-sharded over 4 / 8 GPUs
    One JSON line on rank 0 with the same keys as the headline line."""
    import argparse as _ap
    import importlib
    import torch.distributed as dist
    from repmode_b200 import lib as L, parallel as par, sharded
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = L.load()
    hbm_gbs, tf_burst, tf_sust, peak_kind = peaks()
    mod = importlib.import_module("fnet.nn_modules.RepMode")
    torch.manual_seed(0)
    net = mod.Net(_ap.Namespace(adopted_datasets=list(range(T)), gpu_ids=local_rank)).to(dev)
    stream = torch.cuda.current_stream()
    cfg = args.config
    notes = []
    sharded_cfg = cfg in ("cfg4", "cfg5")
    if sharded_cfg:
        want = 4 if cfg == "cfg4" else 8
        dims = os.environ.get("REPMODE_BENCH_CFG_DIMS")          # dry-run hook: "D,H,W" of a smaller volume on any world
        if world != want and not dims:
            raise SystemExit(f"bench.py --config {cfg} needs --gpus {want} (one rank per GPU under torchrun)")
        Dg, Hg, Wg = (64, 256, 256) if cfg == "cfg4" else (128, 512, 512)
        if dims:
            Dg, Hg, Wg = (int(v) for v in dims.split(","))
            notes.append(f"DRY RUN on a {Dg}x{Hg}x{Wg} volume over {world} GPUs (REPMODE_BENCH_CFG_DIMS): not config {cfg[-1]}")
        dl = Dg // world
        batch = 1
        vox_step = Dg * Hg * Wg
        g = torch.Generator().manual_seed(100 + rank)
        x_host = torch.randn(1, 1, dl, Hg, Wg, generator=g).pin_memory()
        tgt = torch.randn(1, 1, dl, Hg, Wg, generator=g).to(dev)
        task = torch.tensor([3], device=dev)
        flop_step = 3 * NET_FLOP_FWD * vox_step / VOX
    else:
        batch = 1 if cfg == "net_fwd" else 4
        vox_step = batch * VOX
        g = torch.Generator().manual_seed(100)
        x_host = torch.randn(batch, 1, D, H, W, generator=g).pin_memory()
        tgt = torch.randn(batch, 1, D, H, W, generator=g).to(dev)
        task = ((torch.arange(batch) * 3) % T).to(dev) if cfg == "net_train" else torch.tensor([3], device=dev)
        flop_step = (1 if cfg == "net_fwd" else 3) * NET_FLOP_FWD * batch
    x_dev = x_host.to(dev)

    if cfg == "net_fwd":
        net.eval()

        def step(xin):
            with torch.no_grad():
                return net(xin, task)          # from the 2nd call on: one CUDA-graph replay (nn_modules.Net)
        run = lambda: step(x_dev)              # noqa: E731
        result_bytes = vox_step * 4
    else:
        net.train()
        params = list(net.parameters())
        from repmode_b200.optim import FusedAdam           # the path's multi-tensor Adam (csrc/optim.cu): 2 launches per step
        opt = FusedAdam(params, lr=1e-4)
        loss_buf = torch.zeros((), device=dev)

        def train_step(xin):
            opt.zero_grad(set_to_none=False)
            if sharded_cfg:
                out = sharded.sharded_net_forward(net, xin, task, Dg)
                loss = torch.sum((out - tgt) ** 2) / float(vox_step)       # each rank: its slab's share of the global mean
            else:
                out = net(xin, task)
                loss = torch.mean((out - tgt) ** 2)
            loss.backward()
            if world > 1:
                par.sync_gradients(params)
            opt.step()
            loss_buf.copy_(loss.detach())
            return loss_buf
        # warm-up (allocator, lazy state), then try to capture the whole step (fwd + bwd + Adam) as ONE CUDA graph
        for p in params:
            p.grad = torch.zeros_like(p)
        train_step(x_dev)
        l_tr0 = lib.mode_launch_count()
        train_step(x_dev)
        launches_eager_step = lib.mode_launch_count() - l_tr0      # a graph replay launches the same kernels
        torch.cuda.synchronize()
        run = lambda: train_step(x_dev)        # noqa: E731
        # the D-sharded configs go through NCCL collectives (halo send/recv, BatchNorm all-reduces): captured too, with the
        # capture policing this thread only (NCCL's watchdog thread polls events meanwhile); the ranks agree on the outcome
        # so that nobody replays a graph against an eagerly launching neighbour
        want_graph = os.environ.get("REPMODE_BENCH_GRAPH", "1") == "1" and (
            not sharded_cfg or os.environ.get("REPMODE_BENCH_GRAPH_SHARDED", "1") == "1")
        if want_graph:
            graph, why = None, ""
            try:
                if world > 1:
                    dist.barrier()
                graph = torch.cuda.CUDAGraph()
                mode = {"capture_error_mode": "thread_local"} if world > 1 else {}
                with torch.cuda.graph(graph, **mode):
                    train_step(x_dev)
            except Exception as e:  # noqa: BLE001
                graph, why = None, f"{type(e).__name__}: {str(e)[:100]}"
                torch.cuda.synchronize()
            if world > 1:
                ok = torch.tensor([1.0 if graph is not None else 0.0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if float(ok.item()) == 0.0 and graph is not None:
                    graph, why = None, "another rank's capture failed"
            if graph is not None:
                run = graph.replay
                notes.append("train step (fwd + bwd + mode_adam_step) captured once as a CUDA graph and replayed")
            else:
                notes.append(f"CUDA-graph capture failed, eager launches ({why})")
        result_bytes = 4

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
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, lib.mode_launch_count() - l0

    warmup = max(3, args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l_eager0 = lib.mode_launch_count()
    run()                                                               # (net_fwd: the call that captures the graph)
    launches_one = lib.mode_launch_count() - l_eager0
    if cfg != "net_fwd":
        launches_one = max(launches_one, launches_eager_step)
    ms, launches = timed(run, args.steps, warmup)
    clocks = sampler.stop() if sampler else None
    # e2e: input from pinned host memory every step + the step's result read back (prediction / loss)
    x_in = torch.empty_like(x_dev)
    res_host = torch.empty(result_bytes // 4, dtype=torch.float32).pin_memory()

    def e2e_step():
        x_in.copy_(x_host, non_blocking=True)
        if cfg == "net_fwd":
            out = step(x_in)
            res_host.copy_(out.reshape(-1), non_blocking=True)
        else:
            x_dev.copy_(x_in)
            out = run()
            res_host.copy_(loss_buf.reshape(-1), non_blocking=True)
        stream.synchronize()
    ms_e2e, _ = timed(e2e_step, args.steps, warmup)
    if rank != 0:
        return
    tflops = flop_step / (ms / args.steps * 1e-3) / 1e12 / world
    names = {"net_fwd": "full RepMode U-Net eval forward, x[1,1,32,128,128], task 3 (BASELINE.json config 2)",
             "net_train": "full RepMode U-Net train step (fwd + bwd + Adam), x[4,1,32,128,128], 4 distinct tasks of 12 "
                          "(BASELINE.json config 3)",
             "cfg4": "full RepMode U-Net train step on ONE volume x[1,1,64,256,256] sharded on D over 4 GPUs (config 4)",
             "cfg5": "full RepMode U-Net train step on ONE volume x[1,1,128,512,512] sharded on D over 8 GPUs (config 5)"}
    line = {
        "metric": "voxels/sec full RepMode U-Net " + ("eval forward" if cfg == "net_fwd" else "train step"),
        "value": vox_step * args.steps / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if sharded_cfg else "weak", "vs_baseline": None,
        "dtype": "f16 operands (saturating; 10-bit mantissa like TF32), f32 accumulate", "data": "synthetic",
        "config": {"workload": names[cfg], "layout": "NDHWC (channels_last_3d)", "notes": notes,
                   "parallelism": (f"d-shard x{world} (halo exchange per two-conv stage, BatchNorm all-reduce, NCCL)"
                                   if sharded_cfg else "1 GPU")},
        "e2e": {"value": vox_step * args.steps / (ms_e2e * 1e-3), "unit": "voxels/s",
                "h2d_bytes_per_step": x_host.numel() * 4 * world, "d2h_bytes_per_step": result_bytes * world,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches if launches > 0 else launches_one * args.steps),
        "clocks": clocks,
        "roofline": {"kernel": "all MoDE convs of the U-Net (K2 / K3 / K4)", "bound": "tensor", "achieved": tflops,
                     "peak": tf_burst, "unit": "TFLOP/s", "frac": tflops / tf_burst, "traffic": None, "per_gpu": True,
                     "algorithmic_flop_per_step": flop_step,
                     "peak_source": f"MEASURED_PEAKS.json bf16 burst ({peak_kind})"},
        "cpu_baseline": None,
        "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the MoDE-conv path has no CPU fallback); use --impl reference")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.config != "headline":
            run_net_config(args, rank, local_rank, world)
        else:
            run_ours(args, rank, local_rank, world)
        torch.cuda.synchronize()
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
