"""Per-image regression metrics used by Model.do_eval_iter (host-side post-processing; mirrors the
quantities of the reference's fnet/metric.py:7-34 -- MSE, MAE, R^2 on flattened arrays -- with numpy only)."""
import numpy as np


def get_metric_stats(pred, target):
    p = np.asarray(pred.detach().cpu().numpy() if hasattr(pred, "detach") else pred, dtype=np.float64).reshape(-1)
    t = np.asarray(target.detach().cpu().numpy() if hasattr(target, "detach") else target, dtype=np.float64).reshape(-1)
    err = p - t
    mse = float(np.mean(err ** 2))
    mae = float(np.mean(np.abs(err)))
    ss_tot = float(np.sum((t - t.mean()) ** 2))
    r2 = float(1.0 - np.sum(err ** 2) / ss_tot) if ss_tot > 0 else float("nan")
    stats = {"mse": mse, "mae": mae, "r2": r2}
    return (mse, mae, r2), stats
