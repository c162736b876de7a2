"""Oracle (test infrastructure, not product): pointmap -> depth and intrinsics.

Restates the z-extraction sites (scripts/pseudo_gt.py:115-116,
thermal_dustr_inference.py:133-134, utils/metrics.py:121,
utils/evaluate_depth_metrics.py:134-135), the GT nearest resample
(utils/evaluate_depth_metrics.py:320-323) and
scripts/pseudo_gt.py:137-184 (estimate_camera_intrinsics), :232-289
(load_thermal_calibration).  The reference never applies K to a pointmap
(SURVEY.md 8 a-7); `project_points` is the clearly-labelled extension.
"""
from __future__ import annotations

import json

import numpy as np

from .ref_preprocess import resize_nearest


def pointmap_to_depth(pointmap: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(pointmap)[..., 2])


def match_gt(gt_depth: np.ndarray, pred_hw) -> np.ndarray:
    if gt_depth.shape != tuple(pred_hw):
        return resize_nearest(gt_depth, pred_hw)
    return gt_depth


def load_thermal_calibration(path: str):
    """scripts/pseudo_gt.py:232-289."""
    if path.endswith(".json"):
        with open(path) as f:
            c = json.load(f)
        fx, fy, cx, cy = c["intrinsic"]
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        return K, np.array(c["rotation"]), np.array(c["translation"])
    if path.endswith(".yaml"):
        import yaml
        with open(path) as f:
            c = yaml.safe_load(f)
        fx, fy, cx, cy = c["left"]["intrinsics"]
        Kl = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        if "right" in c:
            fx, fy, cx, cy = c["right"]["intrinsics"]
            Kr = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
            return Kl, Kr, np.array(c["right"]["T_cn_cnm1"])
        return Kl, None, None
    raise ValueError(f"Unsupported calibration file format: {path}")


def estimate_camera_intrinsics(pointmap: np.ndarray, depth: np.ndarray) -> np.ndarray:
    """scripts/pseudo_gt.py:151-184 (estimation branch): median focal estimate."""
    H, W = depth.shape
    v, u = np.indices((H, W))
    X, Y, Z = pointmap[:, :, 0], pointmap[:, :, 1], depth
    m = Z > 0
    with np.errstate(all="ignore"):
        xn = X[m] / Z[m]
        yn = Y[m] / Z[m]
        fx = np.median((u[m] - W / 2) / xn)
        fy = np.median((v[m] - H / 2) / yn)
    return np.array([[fx, 0, W / 2], [0, fy, H / 2], [0, 0, 1]])


def project_points(pointmap: np.ndarray, K: np.ndarray):
    """EXTENSION (not in the reference): u = fx X/Z + cx, v = fy Y/Z + cy."""
    X, Y, Z = (pointmap[..., i].astype(np.float32) for i in range(3))
    fx, fy, cx, cy = (np.float32(K[0, 0]), np.float32(K[1, 1]), np.float32(K[0, 2]), np.float32(K[1, 2]))
    with np.errstate(all="ignore"):
        return fx * (X / Z) + cx, fy * (Y / Z) + cy
