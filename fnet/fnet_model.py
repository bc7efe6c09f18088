"""Host wrapper with the reference's `fnet.fnet_model.Model` API (fnet/fnet_model.py:16-239 in the reference
tree) so that main.py / eval.py style drivers run on the B200 network unchanged: same constructor, attributes
(`net, optimizer, count_iter, count_epoch, patch_size, device, gpu_ids`) and methods (`do_train_iter`,
`do_eval_iter`, `predict`, `save_state`, `load_state`, `get_state`, `to_gpu`).  Written from the behaviour, not
from the source.  Deviations: `torch.load(..., weights_only=False)` (the reference call at :85 fails on
torch >= 2.6 because the checkpoint pickles an argparse.Namespace), wandb logging only if wandb is importable,
and `torch.amp` spellings of autocast / GradScaler.
"""
import importlib
import os

import numpy as np
import pandas as pd
import torch

from fnet.metric import get_metric_stats

try:                                     # logging is optional; never on the hot path
    import wandb                         # noqa: F401
except Exception:                        # noqa: BLE001
    wandb = None


def _as_list(gpu_ids):
    return [gpu_ids] if isinstance(gpu_ids, int) else list(gpu_ids)


def _device_of(gpu_ids):
    return torch.device("cuda", gpu_ids[0]) if gpu_ids[0] >= 0 else torch.device("cpu")


class Model(object):
    def __init__(self, opts, nn_module=None, init_weights=True, lr=0.001, criterion_fn=torch.nn.MSELoss, gpu_ids=-1):
        self.opts = opts
        self.nn_module = nn_module
        self.init_weights = init_weights
        self.lr = lr
        self.count_iter = 0
        self.count_epoch = 0
        self.gpu_ids = _as_list(gpu_ids)
        self.device = _device_of(self.gpu_ids)
        self.patch_size = (32, 128, 128)
        self.criterion = criterion_fn(reduction="none")
        self._init_model()
        if self.net is not None and len(self.gpu_ids) > 1:
            raise NotImplementedError("multi-GPU runs use one process per GPU (repmode_b200.parallel), not "
                                      "torch.nn.DataParallel")
        self.scaler = torch.amp.GradScaler("cuda", enabled=self.device.type == "cuda")

    def _init_model(self):
        if self.nn_module is None:
            self.net = None
            return
        self.net = importlib.import_module("fnet.nn_modules." + self.nn_module).Net(self.opts)
        self.net.to(self.device)
        if self.device.type == "cuda":
            # same class hierarchy / state layout as torch.optim.Adam (checkpoints are interchangeable); the update of all
            # 309 tensors is one launch of the path's multi-tensor kernel and GradScaler needs no host sync
            from repmode_b200.optim import FusedAdam
            self.optimizer = FusedAdam(self.net.parameters(), lr=self.lr)
        else:
            self.optimizer = torch.optim.Adam(self.net.parameters(), lr=self.lr)

    # ------------------------------------------------------------------ checkpointing
    _PLAIN_STATE = ("nn_module", "opts", "count_iter", "count_epoch")      # copied as they are; the two state_dicts follow

    def get_state(self):
        state = {key: getattr(self, key) for key in self._PLAIN_STATE}
        state.update(nn_state=self.net.state_dict(), optimizer_state=self.optimizer.state_dict())
        return state

    def to_gpu(self, gpu_ids):
        self.gpu_ids = _as_list(gpu_ids)
        self.device = _device_of(self.gpu_ids)
        self.net.to(self.device)
        _set_gpu_recursive(self.optimizer.state, self.gpu_ids[0])
        if self.device.type == "cuda" and type(self.optimizer) is torch.optim.Adam:
            # built on the CPU (eval.py / load_state construct with gpu_ids=-1 and move afterwards): same state, fused step
            from repmode_b200.optim import FusedAdam
            state = self.optimizer.state_dict()
            self.optimizer = FusedAdam(self.net.parameters(), lr=self.lr)
            self.optimizer.load_state_dict(state)

    def save_state(self, path_save):
        """Checkpoints are written from the CPU copy of the network and optimizer state, then everything moves back."""
        home = self.gpu_ids
        os.makedirs(os.path.dirname(path_save) or ".", exist_ok=True)
        self.to_gpu(-1)
        try:
            torch.save(self.get_state(), path_save)
        finally:
            self.to_gpu(home)

    def load_state(self, path_load, gpu_ids=-1):
        state = torch.load(path_load, weights_only=False)      # the checkpoint pickles an argparse.Namespace
        state["opts"].gpu_ids = gpu_ids
        for key in ("nn_module", "opts"):
            setattr(self, key, state[key])
        self._init_model()                                     # rebuilds net + optimizer from nn_module / opts
        self.net.load_state_dict(state["nn_state"])
        self.optimizer.load_state_dict(state["optimizer_state"])
        for key in ("count_iter", "count_epoch"):
            setattr(self, key, state[key])
        self.to_gpu(gpu_ids)

    # ------------------------------------------------------------------ training / evaluation
    def do_train_iter(self, signal, target, task):
        signal, target, task = signal.to(self.device), target.to(self.device), task.to(self.device)
        self.net.train()
        self.optimizer.zero_grad()
        with torch.amp.autocast("cuda", enabled=self.device.type == "cuda"):
            output = self.net(signal, task)
            loss_nomean = self.criterion(output, target)
            loss = torch.mean(loss_nomean)
        self.scaler.scale(loss).backward()
        self.scaler.step(self.optimizer)
        self.scaler.update()

        per_sample = torch.mean(loss_nomean.detach().float(), dim=(1, 2, 3, 4)).cpu().numpy()
        self._poll_device("do_train_iter")
        task_host = task.cpu().numpy()
        if wandb is not None and getattr(wandb, "run", None) is not None:
            log = {"X-axis/iter": self.count_iter, "loss/iter": float(loss)}
            for i in sorted(set(int(v) for v in task_host)):
                log[f"loss_iter/{self.opts.adopted_datasets[i]}"] = float(per_sample[task_host == i].mean())
            wandb.log(log)
        frame = pd.DataFrame({"dataset": [self.opts.adopted_datasets[int(i)] for i in task_host],
                              "loss": list(per_sample)})
        return output.detach().float().cpu(), frame

    def _poll_device(self, what):
        """The B200 kernels never hang or throw from the device: a timed-out pipeline / an out-of-range task id raises a
        device flag instead.  Read it where the reference's loop synchronises anyway, so a bad step cannot pass silently."""
        if self.device.type == "cuda":
            from repmode_b200 import lib as _mode_lib
            with torch.cuda.device(self.device):
                _mode_lib.poll_error(f"Model.{what}")

    def do_eval_iter(self, signal, target, task, info):
        pred = self.predict(signal, task, self.patch_size)
        stats = get_metric_stats(pred, target)[1]
        row = {"dataset": info["dataset"], "path_czi": info["path_czi"]}
        row.update(stats)                                      # identification columns first, then the metrics
        return pred, pd.DataFrame([row])

    def predict(self, signal, task, patch_size):
        """Sliding-window inference: overlapping patches (stride = half a patch, last patch clamped to the border),
        Gaussian-weighted blending; every batch of patches shares one task (eval uses sample 0's kernel)."""
        signal, task = signal.to(self.device), task.to(self.device)
        self.net.eval()
        size = tuple(signal.shape[-3:])
        from repmode_b200 import predict as _predict
        bs = max(1, int(getattr(self.opts, "batch_size_eval", 1)))
        if self.device.type == "cuda" and signal.shape[0] == 1:
            # B200 path: blend kernels (csrc/predict.cu); with `predict_group` set (a process group whose ranks hold the same
            # volume and weights) the windows are dealt out to the ranks and the accumulators summed once
            gauss = torch.from_numpy(get_gaussian(patch_size))
            pred = _predict.sliding_window_predict(self.net, signal, task, patch_size, bs, gauss,
                                                   group=getattr(self, "predict_group", None)).cpu()
            self._poll_device("predict")
            return pred
        windows = _predict.windows(size, patch_size)
        gauss = torch.from_numpy(get_gaussian(patch_size)).to(self.device)
        pred_sum = torch.zeros(signal.shape, device=self.device)
        weight_sum = torch.zeros(signal.shape, device=self.device)
        for k in range(0, len(windows), bs):
            chunk = windows[k:k + bs]
            batch = torch.cat([signal[:, :, a[0]:a[1], b[0]:b[1], c[0]:c[1]] for a, b, c in chunk], dim=0)
            with torch.no_grad():
                out = self.net(batch, task.expand(len(chunk)))
                if isinstance(out, tuple):
                    out = out[0]
            for j, (a, b, c) in enumerate(chunk):
                g = gauss[:a[1] - a[0], :b[1] - b[0], :c[1] - c[0]]
                pred_sum[:, :, a[0]:a[1], b[0]:b[1], c[0]:c[1]] += out[j:j + 1].float() * g
                weight_sum[:, :, a[0]:a[1], b[0]:b[1], c[0]:c[1]] += g
        return (pred_sum / weight_sum).cpu()

    def __str__(self):
        parts = (("Network", self.nn_module), ("Loss", self.criterion), ("Optimizer", self.optimizer))
        return "".join(f"{name}:\n{value}\n" for name, value in parts)


def get_gaussian(patch_size, sigma_scale=1 / 8):
    """Importance map for patch blending: a unit impulse at the patch centre blurred with sigma = size/8 per axis,
    normalised to max 1, zeros lifted to the smallest positive value."""
    from scipy.ndimage import gaussian_filter
    imp = np.zeros(patch_size)
    imp[tuple(s // 2 for s in patch_size)] = 1
    g = gaussian_filter(imp, [s * sigma_scale for s in patch_size], order=0, mode="constant", cval=0)
    g = (g / g.max()).astype(np.float32)
    g[g == 0] = g[g != 0].min()
    return g


def _set_gpu_recursive(var, gpu_id):
    """Move every tensor nested in dict `var` (optimizer state) to cuda:gpu_id, or to the CPU for -1.  In place."""
    target = torch.device("cpu") if gpu_id == -1 else torch.device("cuda", gpu_id)
    pending = [var]
    while pending:
        node = pending.pop()
        for key, value in node.items():
            if isinstance(value, dict):
                pending.append(value)
            elif torch.is_tensor(value):
                node[key] = value.to(target)
