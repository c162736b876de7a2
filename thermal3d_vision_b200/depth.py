"""Pointmap -> depth and camera intrinsics -- host side (csrc/t3d_metrics.cu).

Mirrors the z-extraction sites of the reference (scripts/pseudo_gt.py:115-116,
thermal_dustr_inference.py:133-134) and scripts/pseudo_gt.py:137-184
(estimate_camera_intrinsics), :232-289 (load_thermal_calibration).  The reference
never applies K to a pointmap; ``project_points`` is a
labelled extension.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib
from .metrics import _as_cuda


@_lib.on_tensor_device
def pointmap_to_depth(pointmap):
    """depth = pointmap[..., 2] as a dense array (same container type/device as the input)."""
    is_np = isinstance(pointmap, np.ndarray)
    src_cuda = (not is_np) and pointmap.is_cuda
    pm = _as_cuda(pointmap, torch.float32).contiguous()
    if pm.shape[-1] != 3:
        raise ValueError("pointmap must be [...,3]")
    depth = torch.empty(pm.shape[:-1], dtype=torch.float32, device=pm.device)
    rc = _lib.lib().t3d_pointmap_to_depth(_lib.ptr(pm), _lib.ptr(depth), depth.numel(), _lib.current_stream_ptr())
    _lib.check(rc, "t3d_pointmap_to_depth")
    if is_np:
        return depth.cpu().numpy()
    return depth if src_cuda else depth.cpu()


def load_thermal_calibration(calib_path):
    """Drop-in for scripts/pseudo_gt.py:232-289 (host-side file parsing)."""
    if calib_path.endswith(".json"):
        with open(calib_path, "r") as f:
            calib = json.load(f)
        fx, fy, cx, cy = calib["intrinsic"]
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        return K, np.array(calib["rotation"]), np.array(calib["translation"])
    elif calib_path.endswith(".yaml"):
        import yaml
        with open(calib_path, "r") as f:
            calib = yaml.safe_load(f)
        fx, fy, cx, cy = calib["left"]["intrinsics"]
        K_left = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        if "right" in calib:
            fx_r, fy_r, cx_r, cy_r = calib["right"]["intrinsics"]
            K_right = np.array([[fx_r, 0, cx_r], [0, fy_r, cy_r], [0, 0, 1]])
            return K_left, K_right, np.array(calib["right"]["T_cn_cnm1"])
        return K_left, None, None
    raise ValueError(f"Unsupported calibration file format: {calib_path}")


@_lib.on_tensor_device
def estimate_camera_intrinsics(pointmap, depth, calib_path=None):
    """Drop-in for scripts/pseudo_gt.py:137-184: K from the calibration file when given,
    else the median-focal estimate (on the GPU, exact float64 medians)."""
    if calib_path and os.path.exists(calib_path):
        try:
            K, _, _ = load_thermal_calibration(calib_path)
            print(f"Loaded camera intrinsics from {calib_path}")
            return K
        except Exception as e:
            print(f"Error loading calibration: {e}, falling back to estimation")
    pm = _as_cuda(pointmap, torch.float32).contiguous()
    dz = _as_cuda(depth, torch.float32).contiguous()
    H, W = dz.shape
    if tuple(pm.shape) != (H, W, 3):
        raise ValueError("pointmap must be [H,W,3] matching depth [H,W]")
    K = torch.empty(1, 9, dtype=torch.float64, device=pm.device)
    rc = _lib.lib().t3d_estimate_focal(_lib.ptr(pm), _lib.ptr(dz), 1, H, W, _lib.ptr(K), _lib.current_stream_ptr())
    _lib.check(rc, "t3d_estimate_focal")
    return K.cpu().numpy().reshape(3, 3)


@_lib.on_tensor_device
def project_points(pointmap, K):
    """EXTENSION (not in the reference): pixel coordinates u = fx X/Z + cx, v = fy Y/Z + cy -> [...,2]."""
    pm = _as_cuda(pointmap, torch.float32).contiguous()
    K = np.asarray(K, np.float64)
    uv = torch.empty(pm.shape[:-1] + (2,), dtype=torch.float32, device=pm.device)
    rc = _lib.lib().t3d_project_points(_lib.ptr(pm), float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]),
                                       _lib.ptr(uv), uv.numel() // 2, _lib.current_stream_ptr())
    _lib.check(rc, "t3d_project_points")
    return uv
