"""Per-image regression metrics used by Model.do_eval_iter (host-side post-processing).  Same contract as the
reference's fnet/metric.py:7-34: `pred` / `target` are CPU tensors of one image; returns `(err_map, stats)` with
`err_map = |pred - target|` as an ndarray carrying two leading singleton axes and `stats` keyed 'MSE', 'MAE', 'R2'
(the column names main.py / eval.py read back from do_eval_iter's DataFrame).  numpy only (no sklearn import)."""
import numpy as np


def _to_numpy(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def get_metric_stats(pred, target):
    pred = _to_numpy(pred)[None, None]
    target = _to_numpy(target)[None, None]
    err_map = np.abs(pred - target)
    p = pred.reshape(-1).astype(np.float64)
    t = target.reshape(-1).astype(np.float64)
    res = float(np.sum((t - p) ** 2))
    tot = float(np.sum((t - t.mean()) ** 2))
    if tot > 0:
        r2 = 1.0 - res / tot
    else:                                   # sklearn's r2_score convention for a constant target
        r2 = 1.0 if res == 0 else 0.0
    all_stats = {
        'MSE': float(np.mean((t - p) ** 2)),
        'MAE': float(np.mean(np.abs(t - p))),
        'R2': float(r2),
    }
    return err_map, all_stats
