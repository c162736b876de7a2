"""Oracle (test infrastructure, not product): thermal preprocessing.

Restates
* /root/reference/utils/preprocessing.py:6-30   enhance_thermal_contrast
* /root/reference/utils/preprocessing.py:32-73  enhance_thermal_fixed_range
* cv2.resize INTER_LINEAR / INTER_NEAREST as called at
  data/dataset_loader.py:242, thermal_dustr_inference.py:52,
  utils/evaluate_depth_metrics.py:189,321-323 (OpenCV's C++ path, IPP off --
  third-party: opencv-python 4.10.0.84 pinned by requirements.txt:109,
  4.13.0 in this image; recipe in SURVEY.md Appendix B)
* np.percentile(method='linear') (third-party: numpy 2.0.2 pinned by
  requirements.txt:94, 2.3.5 here; numpy/lib/_function_base_impl.py
  `_quantile` + `_lerp`).
Pinned by tests/test_oracle_pin.py against live cv2/numpy/reference and by
tests/golden/preprocess_*.npz.
"""
from __future__ import annotations

import numpy as np

GRAY = (0.299, 0.587, 0.114)
FREIBURG_MIN, FREIBURG_MAX = 21800, 25000   # utils/preprocessing.py:53-54


# --------------------------------------------------------------------------- resize
def _linear_taps(src_dim: int, dst_dim: int):
    """Per-output index taps of cv2's INTER_LINEAR (resize.cpp, IPP off).

    f is rounded to fp32 BEFORE floor; coefficients are fp32 (1-f, f)."""
    scale = np.float64(src_dim) / np.float64(dst_dim)
    d = np.arange(dst_dim, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0
    hi = s >= src_dim - 1
    s[hi] = src_dim - 1
    f[hi] = 0
    s1 = np.minimum(s + 1, src_dim - 1)
    return s, s1, (np.float32(1) - f).astype(np.float32), f


def resize_bilinear(src: np.ndarray, dst_hw) -> np.ndarray:
    """cv2.resize(src, (w, h)) for 1-channel uint16 or float32 `src` [H,W]."""
    dh, dw = dst_hw
    sh, sw = src.shape
    x0, x1, cx0, cx1 = _linear_taps(sw, dw)
    y0, y1, cy0, cy1 = _linear_taps(sh, dh)
    S = src.astype(np.float32)
    # horizontal pass: each product rounded separately (no FMA)
    Hr = (S[:, x0] * cx0[None, :]).astype(np.float32) + (S[:, x1] * cx1[None, :]).astype(np.float32)
    Hr = Hr.astype(np.float32)
    out = (Hr[y0, :] * cy0[:, None]).astype(np.float32) + (Hr[y1, :] * cy1[:, None]).astype(np.float32)
    out = out.astype(np.float32)
    if src.dtype == np.uint16:
        return np.clip(np.rint(out), 0, 65535).astype(np.uint16)   # round-half-even + saturate
    return out


def resize_nearest(src: np.ndarray, dst_hw) -> np.ndarray:
    """cv2.resize(..., interpolation=INTER_NEAREST), utils/evaluate_depth_metrics.py:321-323."""
    dh, dw = dst_hw
    sh, sw = src.shape[:2]
    ys = np.minimum(np.floor(np.arange(dh) * (sh / dh)).astype(np.int64), sh - 1)
    xs = np.minimum(np.floor(np.arange(dw) * (sw / dw)).astype(np.int64), sw - 1)
    return src[ys][:, xs]


# --------------------------------------------------------------------------- percentile
def percentile_linear(x: np.ndarray, q: float) -> np.float64:
    """np.percentile(x, q) for a float32 array (method 'linear'), via full sort.

    vi = (n-1)*(q/100) in fp64; a,b fp32 order stats; d=b-a in fp32;
    result fp64 two-sided lerp."""
    flat = np.sort(np.asarray(x, np.float32).ravel())
    n = flat.size
    if np.isnan(flat[-1]):
        return np.float64(np.nan)
    vi = (n - 1) * (np.float64(q) / 100.0)
    k = int(np.floor(vi))
    g = vi - k
    a = flat[k]
    b = flat[min(k + 1, n - 1)]
    d = np.float32(b - a)
    return np.float64(a) + np.float64(d) * g if g < 0.5 else np.float64(b) - np.float64(d) * (1.0 - g)


def histogram_u16(x: np.ndarray) -> np.ndarray:
    """Exact 65 536-bin histogram of integer-valued data (the K2 histogram)."""
    return np.bincount(np.asarray(x).astype(np.int64).ravel(), minlength=65536).astype(np.uint32)


# --------------------------------------------------------------------------- the two public functions
def collapse_channels(t: np.ndarray):
    """utils/preprocessing.py:13-19.  Returns (plane[H,W] fp32, used_gray: bool)."""
    if t.shape[0] == 3:
        if np.allclose(t[0], t[1]) and np.allclose(t[0], t[2]):
            return t[0], False
        return (np.float32(GRAY[0]) * t[0] + np.float32(GRAY[1]) * t[1] + np.float32(GRAY[2]) * t[2]), True
    return t, False


def enhance_thermal_contrast(t: np.ndarray):
    """utils/preprocessing.py:6-30 on a numpy [3,H,W] / [H,W] float32 array.

    Returns (out[3,H,W] float32, p2, p98)."""
    if t is None:
        return None
    t = np.asarray(t, np.float32)
    plane, _ = collapse_channels(t)
    p2 = percentile_linear(plane, 2.0)
    p98 = percentile_linear(plane, 98.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        y = np.clip((plane.astype(np.float64) - p2) / (p98 - p2), 0, 1).astype(np.float32)  # fp64 math, one RNE
    if y.ndim == 2:
        y = np.repeat(y[None], 3, 0)
    return y, p2, p98


def enhance_thermal_fixed_range(t: np.ndarray, normalized: bool = True):
    """utils/preprocessing.py:32-73 (float32 throughout)."""
    if t is None:
        return None
    t = np.asarray(t, np.float32)
    x = t
    if x.ndim == 3:
        if x.shape[0] == 3 and np.allclose(x[0], x[1]) and np.allclose(x[0], x[2]):
            x = x[0]
        elif x.shape[0] == 1:
            x = x[0]
    if normalized:
        x = x * np.float32(65535.0)
    x = np.clip(x, np.float32(FREIBURG_MIN), np.float32(FREIBURG_MAX))
    x = ((x - np.float32(FREIBURG_MIN)) / np.float32(FREIBURG_MAX - FREIBURG_MIN)).astype(np.float32)
    if x.ndim == 2 and t.ndim == 3:
        x = x[None]
        if t.shape[0] == 3:
            x = np.repeat(x, 3, 0)
    return x


# --------------------------------------------------------------------------- full paths
def train_path(raw_u16: np.ndarray, dst_hw):
    """data/dataset_loader.py:237-249 + :110 : u16 -> resize (u16) -> f32 raw counts
    -> 3ch -> enhance_thermal_contrast.  Returns (out[3,h,w], p2, p98, resized_u16)."""
    r = resize_bilinear(raw_u16, dst_hw)
    t = np.repeat(r.astype(np.float32)[None], 3, 0)
    out, p2, p98 = enhance_thermal_contrast(t)
    return out, p2, p98, r


def inference_path(raw_u16: np.ndarray, dst_hw):
    """thermal_dustr_inference.py:25-60: u16 -> /65535 f32 -> resize f32 -> enhance."""
    x = raw_u16.astype(np.float32) / np.float32(65535.0)
    r = resize_bilinear(x, dst_hw)
    t = np.repeat(r[None], 3, 0)
    out, p2, p98 = enhance_thermal_contrast(t)
    return out, p2, p98, r


# --------------------------------------------------------------------------- synthetic raw frames
def make_raw_frames(n: int, seed: int = 0, hw=(512, 640), night_fraction: float = 0.4):
    """Synthetic 16-bit radiometric frames: 'day' clip(N(22800,400)), 'night'
    clip(N(22300,250)) + sparse hot blobs (+1500); all inside the Freiburg window."""
    rng = np.random.default_rng(seed)
    H, W = hw
    out = np.empty((n, H, W), np.uint16)
    for i in range(n):
        night = rng.random() < night_fraction
        if night:
            f = rng.normal(22300, 250, (H, W))
            for _ in range(6):
                cy, cx = rng.integers(0, H), rng.integers(0, W)
                r = int(rng.integers(4, 24))
                f[max(0, cy - r):cy + r, max(0, cx - r):cx + r] += 1500
        else:
            f = rng.normal(22800, 400, (H, W))
        out[i] = np.clip(f, 0, 65535).astype(np.uint16)
    return out
