"""Oracle (test infrastructure, not product): depth-metric reductions.

Restates /root/reference/utils/metrics.py:4-69 (7 metrics),
/root/reference/utils/evaluate_depth_metrics.py:20-80 (3 metrics) and the
accumulator semantics of utils/metrics.py:72-138 (evaluate_thermal_depth).
All arithmetic float32 as in the reference; an extra
float64 variant gives the exact-arithmetic value for tolerance analysis.
"""
from __future__ import annotations

import numpy as np

KEYS7 = ("abs_rel", "sq_rel", "rmse", "rmse_log", "acc_1", "acc_2", "acc_3")


def _mask(gt, mask):
    if mask is None:
        with np.errstate(invalid="ignore"):
            return (gt > 0) & np.isfinite(gt)               # utils/metrics.py:27
    return np.asarray(mask, bool)


def compute_depth_metrics(pred_depth, gt_depth, mask=None, median_scaling=True, dtype=np.float32):
    """utils/metrics.py:4-69.  `dtype=np.float64` gives the exact-arithmetic variant."""
    pred_depth = np.asarray(pred_depth)
    gt_depth = np.asarray(gt_depth)
    m = _mask(gt_depth, mask)
    pred = pred_depth[m].astype(dtype)
    gt = gt_depth[m].astype(dtype)
    if pred.size == 0:                                       # :34-43 (key names differ: quirk 11)
        return {"abs_rel": np.nan, "sq_rel": np.nan, "rmse": np.nan, "rmse_log": np.nan,
                "a1": 0.0, "a2": 0.0, "a3": 0.0}
    with np.errstate(all="ignore"):
        if median_scaling:
            scale = np.median(gt) / np.median(pred)          # :46-48
            pred = pred * scale
        thresh = np.maximum(gt / pred, pred / gt)            # :51
        a1 = (thresh < 1.25).mean()
        a2 = (thresh < 1.25 ** 2).mean()
        a3 = (thresh < 1.25 ** 3).mean()
        abs_rel = np.mean(np.abs(gt - pred) / gt)            # :56-59
        sq_rel = np.mean(((gt - pred) ** 2) / gt)
        rmse = np.sqrt(np.mean((gt - pred) ** 2))
        rmse_log = np.sqrt(np.mean((np.log(gt) - np.log(pred)) ** 2))
    return {"abs_rel": abs_rel, "sq_rel": sq_rel, "rmse": rmse, "rmse_log": rmse_log,
            "acc_1": a1, "acc_2": a2, "acc_3": a3}


def compute_depth_metrics_eval(pred_depth, gt_depth, mask=None, median_scaling=True):
    """utils/evaluate_depth_metrics.py:20-80 (rmse, acc_1.25, acc_1.25^2)."""
    pred_depth = np.asarray(pred_depth)
    gt_depth = np.asarray(gt_depth)
    m = _mask(gt_depth, mask)
    pred = pred_depth[m].astype(np.float32)
    gt = gt_depth[m].astype(np.float32)
    if pred.size == 0:
        return {"rmse": np.nan, "acc_1.25": 0.0, "acc_1.25^2": 0.0}
    with np.errstate(all="ignore"):
        if median_scaling:
            pred = pred * (np.median(gt) / np.median(pred))
        rmse = np.sqrt(np.mean((gt - pred) ** 2))
        thresh = np.maximum(gt / pred, pred / gt)
        return {"rmse": rmse, "acc_1.25": (thresh < 1.25).mean(), "acc_1.25^2": (thresh < 1.25 ** 2).mean()}


def accumulate_dataset(per_image_metrics):
    """utils/metrics.py:86-136: sum finite metrics, divide by the count of ALL samples."""
    sums = {k: 0.0 for k in KEYS7}
    n = 0
    for m in per_image_metrics:
        for k in KEYS7:
            if np.isfinite(m[k]):
                sums[k] += m[k]
        n += 1
    return {k: (v / n if n > 0 else np.nan) for k, v in sums.items()}
