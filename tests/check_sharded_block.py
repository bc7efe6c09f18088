"""Multi-rank parity check of the D-sharded MoDE-conv block (repmode_b200.sharded.sharded_mode_conv) against the UNSHARDED
oracle (oracle/mode_torch.py on the whole volume, CPU fp32): output slab, dx slab, every parameter gradient (summed over the
slabs) and the BatchNorm running statistics.  Launch under torchrun with one rank per process:

    torchrun --nproc-per-node 2 tests/check_sharded_block.py --comm peer|nccl|gloo [--same-device]

  --comm peer   stores into the neighbour's memory over CUDA IPC (the product path)
  --comm nccl   torch.distributed / NCCL (the baseline; needs one GPU per rank)
  --comm gloo   torch.distributed / gloo with host staging (any number of ranks on ONE GPU: the 1-GPU test box)
  --same-device every rank uses cuda:0 (two processes time-slice one GPU)
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mode_torch as otc  # noqa: E402
from repmode_b200 import peer, sharded  # noqa: E402
from repmode_b200.nn_modules import MoDEConv  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def grad_err(a, b):
    """Gradients are judged by relative L2 (with max-rel kept as a sanity bound): a voxel whose pre-activation sits within
    rounding of zero can take the other side of the ReLU than the CPU oracle (different summation order), and ONE such flip
    moves a per-channel sum by a whole |dout| -- seen once in r2s on 4 GPUs: max-rel 4e-3 on the fp32 path with every other
    entry at 1e-6 (dgamma untouched, because xhat = 0 exactly there)."""
    m, l2 = rel(a, b), rel_l2(a, b)
    return l2 if m <= 0.1 else m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl", "gloo"])
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if args.same_device else int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.comm == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    ok = True
    for precision, tol_out, tol_grad in (("f32", 1e-4, 5e-4), ("f16", 1e-3, 2e-2)):
        torch.manual_seed(0)
        C, Dl, H, W, T = 32, 6, 32, 16, 12
        D = Dl * world
        m = MoDEConv(5, T, C, C).to(dev).train()
        m.precision = precision
        g = torch.Generator().manual_seed(1)
        x = torch.randn(1, C, D, H, W, generator=g)
        dout = torch.randn(1, C, D, H, W, generator=g)
        t = torch.tensor([5])
        # ReLU kinks: a voxel whose pre-activation sits within rounding of zero can take either side of the ReLU depending
        # on the summation order, and ONE such flip moves every per-channel gradient sum by a whole |dout| (r2s / r2t on
        # 4 GPUs, fp32 path: dbeta off by 1e-3 with dgamma exact, i.e. exactly one voxel with xhat ~ 0).  The gradient is not
        # defined there, so those voxels get dout = 0 -- for the oracle and for the GPU run alike.
        with torch.no_grad():
            pd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
            pre = otc.mode_conv(pd, "", x, t, True, conv_type="conv_only", operand_f16=(precision == "f16"))
            z = torch.nn.functional.batch_norm(pre, None, None, pd["subsequent_layer.0.weight"], pd["subsequent_layer.0.bias"],
                                               True, 0.0, 1e-5)
            kink = z.abs() < 1e-4
            dout = dout.masked_fill(kink, 0.0)
        # unsharded oracle on the CPU (same operand rounding as the tensor-core path for precision f16)
        p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        for k in list(p):
            if p[k].dtype.is_floating_point and "running" not in k and "pool" not in k:
                p[k].requires_grad_(True)
        xc = x.clone().requires_grad_(True)
        yc = otc.mode_conv(p, "", xc, t, True, update_running=True, operand_f16=(precision == "f16"))
        yc.backward(dout)

        if args.comm == "peer":
            comm = peer.PeerComm(dev, 64 << 20)
        else:
            comm = peer.TorchComm(stage_host=(args.comm == "gloo"))
        lo = rank * Dl
        errs = {}
        for step in range(args.steps):                           # >1 step: buffers / counters are reused
            for q in m.parameters():
                q.grad = None
            if step > 0:
                m.subsequent_layer[0].reset_running_stats()
            xl = x[:, :, lo:lo + Dl].to(dev).requires_grad_(True)
            yl = sharded.sharded_mode_conv(m, xl, t.to(dev), comm, D, tag="chk")
            yl.backward(dout[:, :, lo:lo + Dl].to(dev))
            torch.cuda.synchronize()
            errs = {"out": (rel(yl.detach().cpu(), yc.detach()[:, :, lo:lo + Dl]), tol_out),
                    "dx": (grad_err(xl.grad.cpu(), xc.grad[:, :, lo:lo + Dl]), tol_grad),
                    "running_mean": (rel(m.subsequent_layer[0].running_mean.cpu(), p["subsequent_layer.0.running_mean"]), tol_out),
                    "running_var": (rel(m.subsequent_layer[0].running_var.cpu(), p["subsequent_layer.0.running_var"]), tol_out)}
            for k, q in m.named_parameters():
                errs[k] = (grad_err(q.grad.cpu(), p[k].grad), tol_grad)
        import ctypes
        from repmode_b200 import lib as L
        code = ctypes.c_int32(0)
        L.check(L.load().mode_poll_error(ctypes.byref(code)), "mode_poll_error")
        bad = {k: v for k, (v, tol) in errs.items() if not v <= tol}
        if code.value != 0:
            bad["device_error_flag"] = code.value
        flag = torch.tensor([1.0 if bad else 0.0], device=dev if args.comm == "nccl" else "cpu")
        dist.all_reduce(flag)
        if rank == 0:
            print(f"[{args.comm} x{world} {precision}] collectives/step={comm.n_collectives // args.steps} kinks={int(kink.sum())} " +
                  " ".join(f"{k}={v:.2e}" for k, (v, _) in errs.items()), flush=True)
        if bad:
            print(f"rank {rank} [{precision}] FAILED: {bad}", flush=True)
        if flag.item() > 0:
            ok = False
        if hasattr(comm, "close"):
            comm.close()
        del comm
    dist.barrier()
    dist.destroy_process_group()
    if ok and rank == 0:
        print("SHARDED_BLOCK_OK", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
