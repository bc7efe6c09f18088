"""Multi-GPU check (run under torchrun, one rank per GPU): the D-sharded two-conv stage (repmode_b200/sharded.py)
against the same stage run unsharded on one GPU, forward and backward (dx and all parameter gradients after the
data-parallel gradient sum).   torchrun --nproc-per-node 2 tests/check_sharded.py"""
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
    # ---- unit checks of the sharded stride-2 levels and of the gather / re-slice transitions (fp32)
    from repmode_b200 import functional as Fm
    torch.manual_seed(2)
    N, C, D, H, W = 1, 16, 8 * world, 16, 16
    dl = D // world
    lo = rank * dl
    conv = torch.nn.Conv3d(C, C, 2, stride=2, bias=False).cuda()
    convt = torch.nn.ConvTranspose3d(C, C // 2, 2, stride=2, bias=False).cuda()
    bn1, bn2 = torch.nn.BatchNorm3d(C).cuda().train(), torch.nn.BatchNorm3d(C // 2).cuda().train()
    with torch.no_grad():
        for bn in (bn1, bn2):
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.5, 0.5)
    x = torch.randn(N, C, D, H, W, device="cuda")
    for name, fn, w_, bn, oc in (("down", Fm.down_conv_bn_relu, conv.weight, bn1, C),
                                 ("up", Fm.up_conv_bn_relu, convt.weight, bn2, C // 2)):
        f = 0.5 if name == "down" else 2.0
        go = torch.randn(N, oc, int(D * f), int(H * f), int(W * f), device="cuda")
        xr = x.clone().requires_grad_(True)
        for p_ in (w_, bn.weight, bn.bias):
            p_.grad = None
        yr = fn(xr, w_, bn, True)
        yr.backward(go)
        ref = [xr.grad.clone(), w_.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()]
        for p_ in (w_, bn.weight, bn.bias):
            p_.grad = None
        xl = x[:, :, lo:lo + dl].clone().requires_grad_(True)
        spec = sharded._full_spec(N, int(dl * f), int(H * f), int(W * f), int(D * f), None)
        yl = fn(xl, w_, bn, True, spec)
        ol, oh = int(lo * f), int((lo + dl) * f)
        yl.backward(go[:, :, ol:oh])
        par.sync_gradients([w_, bn.weight, bn.bias])
        e = [rel(yl, yr[:, :, ol:oh]), rel(xl.grad, ref[0][:, :, lo:lo + dl]), rel(w_.grad, ref[1]),
             rel(bn.weight.grad, ref[2]), rel(bn.bias.grad, ref[3])]
        if rank == 0:
            print(f"[unit {name}] out {e[0]:.2e} dx {e[1]:.2e} dW {e[2]:.2e} dgamma {e[3]:.2e} dbeta {e[4]:.2e}", flush=True)
    # gather -> replicated op -> re-slice
    lin = torch.nn.Conv3d(C, C, 1, bias=False).cuda()
    xr = x.clone().requires_grad_(True)
    go = torch.randn(N, C, D, H, W, device="cuda")
    yr = torch.relu(lin(xr)); yr.backward(go)
    ref_dx, ref_dw = xr.grad.clone(), lin.weight.grad.clone()
    lin.weight.grad = None
    xl = x[:, :, lo:lo + dl].clone().requires_grad_(True)
    yl = torch.relu(lin(sharded._AllGatherD.apply(xl, None)))[:, :, lo:lo + dl]
    yl.backward(go[:, :, lo:lo + dl])
    par.sync_gradients([lin.weight])
    if rank == 0:
        print(f"[unit gather] out {rel(yl, yr[:, :, lo:lo + dl]):.2e} dx {rel(xl.grad, ref_dx[:, :, lo:lo + dl]):.2e} "
              f"dW {rel(lin.weight.grad, ref_dw):.2e}", flush=True)

    # ---- whole U-Net, D-sharded (levels thinner than the halo are replicated), forward + backward
    failures = 0
    import argparse
    from repmode_b200.nn_modules import Net
    torch.backends.cuda.matmul.allow_tf32 = False
    for precision, tol in (("f32", 5e-4), ("f32-torchbn", 5e-4), ("f16", 5e-3)):
        os.environ["REPMODE_SHARD_DEBUG"] = "1" if precision == "f32-torchbn" else "0"
        precision = precision.split("-")[0]
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
        pr_ref, pr_sh = {}, {}
        yr = sharded.sharded_net_forward(net, x, t, D, probe=pr_ref, replicated=True)
        yr.backward(dout)
        ref = {k: p.grad.clone() for k, p in net.named_parameters()}
        net.load_state_dict(sd0)
        for p in net.parameters():
            p.grad = None
        dl = D // world
        Fm.COLL_LOG = []
        yl = sharded.sharded_net_forward(net, x[:, :, rank * dl:(rank + 1) * dl].contiguous(), t, D, probe=pr_sh)
        nfwd = len(Fm.COLL_LOG)
        yl.backward(dout[:, :, rank * dl:(rank + 1) * dl])
        log = Fm.COLL_LOG
        Fm.COLL_LOG = None
        gathered = [None] * world
        dist.all_gather_object(gathered, log)
        if rank == 0:
            same = all(g == gathered[0] for g in gathered)
            print(f"[net {precision} collectives] fwd {nfwd} total {len(log)} identical_across_ranks {same} "
                  f"bwd_head {log[nfwd:nfwd + 8]}", flush=True)
        if rank == 0:
            msgs = []
            for k in pr_ref:
                a, b = pr_sh[k], pr_ref[k]
                if a.shape[2] == b.shape[2]:            # replicated level: gradients are partial sums -> skip grads
                    msgs.append(f"{k}: act {rel(a, b):.1e} (replicated)")
                else:
                    w_ = a.shape[2]
                    msgs.append(f"{k}: act {rel(a, b[:, :, rank * w_:(rank + 1) * w_]):.1e} "
                                f"grad {rel(a.grad, b.grad[:, :, rank * w_:(rank + 1) * w_]):.1e}")
            print(f"[net {precision} probe] " + " | ".join(msgs), flush=True)
        par.sync_gradients(list(net.parameters()))
        # arbiter: the torch restatement of the reference (oracle/mode_torch.py) with autograd, fp32, no TF32
        from oracle import mode_torch as orc
        torch.backends.cudnn.allow_tf32 = False
        po = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and k in dict(net.named_parameters())
                  else v.clone()) for k, v in sd0.items()}
        yo = orc.net_forward(po, x, t.to(torch.int64), True)
        yo.backward(dout)
        e_ref = {k: rel(ref[k], po[k].grad) for k, _ in net.named_parameters()}
        e_sh = {k: rel(pp.grad, po[k].grad) for k, pp in net.named_parameters()}
        if rank == 0:
            for nm, e in (("unsharded-vs-oracle", e_ref), ("sharded-vs-oracle", e_sh)):
                top = sorted(e.items(), key=lambda kv: -kv[1])[:4]
                print(f"[net {precision} {nm}] out {rel(yr, yo):.1e} n_bad {sum(v > tol for v in e.values())}/{len(e)} worst: "
                      + ", ".join(f"{k} {v:.1e}" for k, v in top), flush=True)
        errs = {"out": rel(yl, yr[:, :, rank * dl:(rank + 1) * dl])}
        for k, p in net.named_parameters():
            errs[k] = rel(p.grad, ref[k])
        bad = {k: v for k, v in errs.items() if not (v <= tol)}
        if rank == 0:
            top = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
            print(f"[net {precision}] world={world} out {errs['out']:.2e}  n_bad {len(bad)}/{len(errs)}  worst: "
                  + ", ".join(f"{k} {v:.1e}" for k, v in top), flush=True)
        if rank == 0 and bad:
            print(f"[net {precision}] FAILED with {len(bad)} tensors over tolerance", flush=True)
        failures += 1 if bad else 0
    dist.barrier()
    assert failures == 0, f"{failures} whole-net configurations failed"
    if rank == 0:
        print("SHARDED_CHECK_OK", worst, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
