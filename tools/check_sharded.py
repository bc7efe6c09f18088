"""Multi-GPU check (run under torchrun, one rank per GPU): the D-sharded two-conv stage (repmode_b200/sharded.py)
against the same stage run unsharded on one GPU, forward and backward (dx and all parameter gradients after the
data-parallel gradient sum).   torchrun --nproc-per-node 2 tools/check_sharded.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from repmode_b200 import parallel as par, sharded  # noqa: E402
from repmode_b200.nn_modules import MoDESubNet2Conv  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    worst = 0.0
    for precision, tol in (("f32", 2e-4), ("f16", 2e-3)):
        torch.manual_seed(0)
        N, C, D, H, W, T = 1, 32, 8 * world, 32, 16, 12
        stage = MoDESubNet2Conv(5, T, C, C).cuda().train()
        for m in (stage.conv1, stage.conv2):
            m.precision = precision
        x = torch.randn(N, C, D, H, W, device="cuda")
        dout = torch.randn(N, C, D, H, W, device="cuda")
        t = torch.tensor([5], device="cuda", dtype=torch.int32)
        # unsharded reference run (every rank computes it)
        xr = x.clone().requires_grad_(True)
        yr = stage(xr, t)
        yr.backward(dout)
        ref_grads = {k: p.grad.clone() for k, p in stage.named_parameters()}
        ref_dx = xr.grad.clone()
        ref_rm = stage.conv1.subsequent_layer[0].running_mean.clone()
        for p in stage.parameters():
            p.grad = None
        for m in (stage.conv1, stage.conv2):          # undo the reference run's running-stat update
            m.subsequent_layer[0].reset_running_stats()
        # sharded run
        dl = D // world
        lo = rank * dl
        xl = x[:, :, lo:lo + dl].clone().requires_grad_(True)
        yl = sharded.sharded_stage(stage, xl, t, D)
        yl.backward(dout[:, :, lo:lo + dl])
        par.sync_gradients(list(stage.parameters()))
        errs = {"out": rel(yl, yr[:, :, lo:lo + dl]), "dx": rel(xl.grad, ref_dx[:, :, lo:lo + dl]),
                "running_mean": rel(stage.conv1.subsequent_layer[0].running_mean, ref_rm)}
        for k, p in stage.named_parameters():
            errs[k] = rel(p.grad, ref_grads[k])
        bad = {k: v for k, v in errs.items() if not (v <= tol)}
        worst = max(worst, max(errs.values()))
        if rank == 0:
            print(f"[{precision}] world={world} max rel err {max(errs.values()):.3e}  (out {errs['out']:.2e}, dx {errs['dx']:.2e})",
                  flush=True)
        assert not bad, (precision, bad)
    # ---- whole U-Net, D-sharded (levels thinner than the halo are replicated), forward + backward
    import argparse
    from repmode_b200.nn_modules import Net
    torch.backends.cuda.matmul.allow_tf32 = False
    for precision, tol in (("f32", 5e-4), ("f16", 5e-3)):
        torch.manual_seed(1)
        net = Net(argparse.Namespace(adopted_datasets=list(range(12)), gpu_ids=local), mult_chan=8).cuda().train()
        for m in net.modules():
            if hasattr(m, "precision"):
                m.precision = precision
        D, H, W = 16 * world, 32, 32
        x = torch.randn(1, 1, D, H, W, device="cuda")
        dout = torch.randn(1, 1, D, H, W, device="cuda")
        t = torch.tensor([4], device="cuda")
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        yr = net(x, t)
        yr.backward(dout)
        ref = {k: p.grad.clone() for k, p in net.named_parameters()}
        net.load_state_dict(sd0)
        for p in net.parameters():
            p.grad = None
        dl = D // world
        yl = sharded.sharded_net_forward(net, x[:, :, rank * dl:(rank + 1) * dl].contiguous(), t, D)
        yl.backward(dout[:, :, rank * dl:(rank + 1) * dl])
        par.sync_gradients(list(net.parameters()))
        errs = {"out": rel(yl, yr[:, :, rank * dl:(rank + 1) * dl])}
        for k, p in net.named_parameters():
            errs[k] = rel(p.grad, ref[k])
        bad = {k: v for k, v in errs.items() if not (v <= tol)}
        worst_k = max(errs, key=errs.get)
        if rank == 0:
            print(f"[net {precision}] world={world} out {errs['out']:.2e}  worst grad {worst_k} {errs[worst_k]:.2e}", flush=True)
        assert not bad, (precision, bad)
    dist.barrier()
    if rank == 0:
        print("SHARDED_CHECK_OK", worst, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
