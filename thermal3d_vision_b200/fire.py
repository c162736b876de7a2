"""Experimental fire-scene pipeline -- host side of csrc/t3d_fire.cu.

Mirrors /root/reference/thermal_dustr_inference_for_experiment.py:62-377 (same function names, arguments and return
conventions):

    preprocess_fire_scene_thermal(thermal_img)                     :62-152
    advanced_fire_scene_processing(thermal_img)                    :154-282
    depth_refinement_with_outlier_removal(depth_map, thermal_img, guided_filter=True)   :284-377

The reference strings OpenCV / NumPy / SciPy calls together on the CPU; here every image operator (CLAHE, Canny,
Sobel, bilateral filter, percentiles, 100-bin histogram, outlier median, the per-pixel compositions) is a kernel of
libt3d_sm100.so.  What stays on the host is control logic on a handful of numbers: the peak search over the 100
histogram counts (scipy.signal.find_peaks(height=, distance=) restated below) and the `np.random.rand(h, w)` texture,
which is drawn from NumPy's global generator at the same point as in the reference so that a seeded run reproduces
the reference's output.  `cv2.ximgproc.guidedFilter` (:358-370) is not available in this environment and has no
oracle: `guided_filter=True` raises.  There is no CPU implementation.

Device rule as elsewhere: CUDA tensor in -> CUDA tensor out; CPU tensor / ndarray in -> CPU tensor out (the
reference's return type), `depth_refinement_with_outlier_removal` -> ndarray like the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _cuda_device():
    if not torch.cuda.is_available():
        raise _lib.T3DError("no CUDA device: thermal3d_vision_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _to_chw_cuda(thermal_img):
    """The reference's input normalisation (:71-92): -> (CUDA float32 [C,H,W] or [H,W], was_cuda)."""
    if isinstance(thermal_img, torch.Tensor):
        was_cuda = thermal_img.is_cuda
        t = thermal_img.detach()
        hwc = False                                            # a [3,H,W] tensor is transposed to HWC there: channels first here
        if not (t.dim() == 3 and t.shape[0] == 3) and t.dim() == 3:
            hwc = True                                         # any other 3-D tensor is taken as [H,W,C]
    else:
        was_cuda = False
        t = torch.from_numpy(np.ascontiguousarray(np.array(thermal_img)))
        hwc = t.dim() == 3
    dev = t.device if t.is_cuda else _cuda_device()
    t = t.to(dev)
    if t.dtype != torch.float32:                               # :80-83
        t = t.to(torch.float32)
        if t.numel() and float(t.max()) > 1.0:
            t = t / 255.0
    if t.dim() == 3 and hwc:
        if t.shape[2] >= 3:
            t = t[:, :, :3].permute(2, 0, 1)
        elif t.shape[2] == 1:
            t = t[:, :, 0]
        else:
            raise ValueError(f"unsupported thermal image shape {tuple(t.shape)}")
    if t.dim() not in (2, 3):
        raise ValueError(f"unsupported thermal image shape {tuple(t.shape)}")
    return t.contiguous(), was_cuda


def _gray(t):
    """:86-92 -> float32 [H,W] on the device."""
    lib = _lib.lib()
    if t.dim() == 2:
        return t
    C, H, W = t.shape
    g = torch.empty(H, W, dtype=torch.float32, device=t.device)
    _lib.check(lib.t3d_fire_gray(_lib.ptr(t), int(C), H, W, _lib.ptr(g), _lib.current_stream_ptr()), "t3d_fire_gray")
    return g


def _clahe(u8, clip):
    lib = _lib.lib()
    H, W = u8.shape
    ws = torch.empty(lib.t3d_clahe_workspace_bytes(1, 8, 8), dtype=torch.uint8, device=u8.device)
    out = torch.empty_like(u8)
    _lib.check(lib.t3d_clahe_u8(_lib.ptr(u8), _lib.ptr(out), 1, H, W, float(clip), 8, 8, _lib.ptr(ws), ws.numel(),
                                _lib.current_stream_ptr()), "t3d_clahe_u8")
    return out


def _canny(u8, low, high):
    lib = _lib.lib()
    H, W = u8.shape
    ws = torch.empty(lib.t3d_canny_workspace_bytes(H, W), dtype=torch.uint8, device=u8.device)
    out = torch.empty_like(u8)
    _lib.check(lib.t3d_canny_u8(_lib.ptr(u8), _lib.ptr(out), H, W, float(low), float(high), _lib.ptr(ws), ws.numel(),
                                _lib.current_stream_ptr()), "t3d_canny_u8")
    return out


def _bilateral(x_hwc, d, sigma_color, sigma_space):
    lib = _lib.lib()
    H, W = x_hwc.shape[:2]
    C = 1 if x_hwc.dim() == 2 else int(x_hwc.shape[2])
    out = torch.empty_like(x_hwc)
    scratch = torch.empty(2, dtype=torch.float32, device=x_hwc.device)
    _lib.check(lib.t3d_bilateral_f32(_lib.ptr(x_hwc), _lib.ptr(out), H, W, C, int(d), float(sigma_color), float(sigma_space),
                                     _lib.ptr(scratch), _lib.current_stream_ptr()), "t3d_bilateral_f32")
    return out


def _noise(h, w, dev):
    """`np.random.rand(h, w).astype(np.float32)` (:126, :256): the same draw from NumPy's global generator."""
    return torch.from_numpy(np.random.rand(h, w).astype(np.float32)).to(dev)


# scipy.signal.find_peaks(x, height=h, distance=d)[0] on the 100 histogram counts (host control logic)
def _find_peaks(x, height, distance):
    x = np.asarray(x, np.float64)
    n = len(x)
    mids = []
    i = 1
    while i < n - 1:                                           # _local_maxima_1d: plateaus -> their midpoint
        if x[i - 1] < x[i]:
            a = i + 1
            while a < n - 1 and x[a] == x[i]:
                a += 1
            if x[a] < x[i]:
                mids.append((i + a - 1) // 2)
                i = a
        i += 1
    peaks = np.array([q for q in mids if x[q] >= height], np.intp)
    if len(peaks):                                             # _select_by_peak_distance: highest first
        keep = np.ones(len(peaks), bool)
        order = np.argsort(x[peaks])
        for t in range(len(peaks) - 1, -1, -1):
            j = order[t]
            if not keep[j]:
                continue
            k = j - 1
            while k >= 0 and peaks[j] - peaks[k] < distance:
                keep[k] = False
                k -= 1
            k = j + 1
            while k < len(peaks) and peaks[k] - peaks[j] < distance:
                keep[k] = False
                k += 1
        peaks = peaks[keep]
    return peaks


@_lib.on_tensor_device
def preprocess_fire_scene_thermal(thermal_img):
    """Drop-in for thermal_dustr_inference_for_experiment.py:62-152 -> float32 [3,H,W]."""
    lib = _lib.lib()
    t, was_cuda = _to_chw_cuda(thermal_img)
    with _lib.device_guard(t.device):
        g = _gray(t)
        H, W = g.shape
        dev, st = g.device, _lib.current_stream_ptr()
        pct = torch.empty(1, 2, dtype=torch.float64, device=dev)
        ws = torch.empty(lib.t3d_contrast_normalize_workspace_bytes(1), dtype=torch.uint8, device=dev)
        _lib.check(lib.t3d_percentiles_f32(_lib.ptr(g), 1, H * W, 5.0, 95.0, _lib.ptr(pct), _lib.ptr(ws), ws.numel(), st),
                   "t3d_percentiles_f32")                       # :95
        base_u8 = torch.empty(H, W, dtype=torch.uint8, device=dev)
        norm_u8 = torch.empty(H, W, dtype=torch.uint8, device=dev)
        _lib.check(lib.t3d_fire_norm_u8(_lib.ptr(g), H, W, _lib.ptr(pct), _lib.ptr(base_u8), _lib.ptr(norm_u8), st), "t3d_fire_norm_u8")
        clahe = _clahe(base_u8, 3.0)                            # :108-109
        noise = _noise(H, W, dev)                               # :126
        edges = _canny(norm_u8, 50, 150)                        # :135
        out = torch.empty(3, H, W, dtype=torch.float32, device=dev)
        _lib.check(lib.t3d_fire_compose(_lib.ptr(g), H, W, _lib.ptr(pct), _lib.ptr(clahe), _lib.ptr(edges), _lib.ptr(noise),
                                        _lib.ptr(out), st), "t3d_fire_compose")
    return out if was_cuda else out.cpu()


@_lib.on_tensor_device
def advanced_fire_scene_processing(thermal_img):
    """Drop-in for thermal_dustr_inference_for_experiment.py:154-282 -> float32 [3,H,W]."""
    lib = _lib.lib()
    t, was_cuda = _to_chw_cuda(thermal_img)
    with _lib.device_guard(t.device):
        g = _gray(t)
        H, W = g.shape
        dev, st = g.device, _lib.current_stream_ptr()
        hist = torch.empty(100, dtype=torch.int32, device=dev)
        _lib.check(lib.t3d_histogram100(_lib.ptr(g), H * W, _lib.ptr(hist), st), "t3d_histogram100")       # :188
        h = hist.cpu().numpy().view(np.uint32).astype(np.int64)                                             # 400 bytes D2H
        bins = np.linspace(0.0, 1.0, 101)
        peaks = _find_peaks(h, h.max() * 0.3, 10)                                                           # :192
        peak_values = np.sort(bins[peaks])
        # :196-214: only the last mask (the hottest region) is used below -- its lower bound
        fire_threshold = float((peak_values[-2] + peak_values[-1]) / 2) if len(peak_values) >= 2 else 0.7
        inv_u8 = torch.empty(H, W, dtype=torch.uint8, device=dev)
        gray_u8 = torch.empty(H, W, dtype=torch.uint8, device=dev)
        _lib.check(lib.t3d_fire_adv_u8(_lib.ptr(g), H, W, _lib.ptr(inv_u8), _lib.ptr(gray_u8), st), "t3d_fire_adv_u8")
        clahe = _clahe(inv_u8, 2.5)                             # :220-221
        edges1 = _canny(gray_u8, 30, 150)                       # :225
        noise = _noise(H, W, dev)                               # :256
        hwc = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
        scratch = torch.empty(2 * H * W + 2, dtype=torch.float32, device=dev)
        _lib.check(lib.t3d_fire_adv_compose(_lib.ptr(g), H, W, fire_threshold, _lib.ptr(clahe), _lib.ptr(edges1), _lib.ptr(noise),
                                            _lib.ptr(hwc), _lib.ptr(scratch), st), "t3d_fire_adv_compose")
        filt = _bilateral(hwc, 9, 75, 75)                       # :273
        out = torch.empty(3, H, W, dtype=torch.float32, device=dev)
        _lib.check(lib.t3d_hwc_to_chw_clip01(_lib.ptr(filt), H, W, _lib.ptr(out), st), "t3d_hwc_to_chw_clip01")
    return out if was_cuda else out.cpu()


@_lib.on_tensor_device
def depth_refinement_with_outlier_removal(depth_map, thermal_img, guided_filter=True):
    """Drop-in for thermal_dustr_inference_for_experiment.py:284-377 without the guided filter: 3-sigma outliers ->
    median of their 5x5 inlier neighbours (:335-356), then cv2.bilateralFilter(., 5, 50, 50) (:375).

    `guided_filter=True` (the reference's default) needs cv2.ximgproc.guidedFilter, which this environment does not
    have (no oracle): it raises instead of silently skipping the step.  Returns an ndarray like the reference
    (a CUDA tensor when `depth_map` is one)."""
    lib = _lib.lib()
    was_cuda = isinstance(depth_map, torch.Tensor) and depth_map.is_cuda
    d = depth_map.detach() if isinstance(depth_map, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(depth_map)))
    if d.dim() != 2:
        raise ValueError(f"expected an [H,W] depth map, got {tuple(d.shape)}")
    if guided_filter:
        th_shape = tuple(thermal_img.shape[-2:]) if hasattr(thermal_img, "shape") and len(thermal_img.shape) >= 2 else None
        if th_shape == tuple(d.shape):                          # :359: the filter only runs when the shapes match
            raise NotImplementedError("guided_filter=True needs cv2.ximgproc.guidedFilter (opencv-contrib), which has no "
                                      "oracle in this environment; pass guided_filter=False")
    dev = d.device if d.is_cuda else _cuda_device()
    d = d.to(dev, torch.float32).contiguous()
    with _lib.device_guard(dev):
        H, W = d.shape
        st = _lib.current_stream_ptr()
        cleaned = torch.empty_like(d)
        stats = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.t3d_depth_outlier_median(_lib.ptr(d), _lib.ptr(cleaned), H, W, _lib.ptr(stats), None, st),
                   "t3d_depth_outlier_median")
        out = _bilateral(cleaned, 5, 50, 50)
    return out if was_cuda else out.cpu().numpy()


# the individual operators, for callers that want them (all take / return CUDA tensors)
@_lib.on_tensor_device
def clahe_u8(img_u8: torch.Tensor, clip_limit: float, tiles=(8, 8)) -> torch.Tensor:
    """cv2.createCLAHE(clipLimit, tiles).apply(img) for uint8 [H,W] or [B,H,W] CUDA tensors."""
    lib = _lib.lib()
    _lib.require_cuda(img_u8)
    x = img_u8.contiguous()
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(0)
    B, H, W = x.shape
    ws = torch.empty(lib.t3d_clahe_workspace_bytes(B, tiles[0], tiles[1]), dtype=torch.uint8, device=x.device)
    out = torch.empty_like(x)
    _lib.check(lib.t3d_clahe_u8(_lib.ptr(x), _lib.ptr(out), B, H, W, float(clip_limit), int(tiles[0]), int(tiles[1]),
                                _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr()), "t3d_clahe_u8")
    return out[0] if squeeze else out


@_lib.on_tensor_device
def canny_u8(img_u8: torch.Tensor, low: float, high: float) -> torch.Tensor:
    """cv2.Canny(img, low, high) for a uint8 [H,W] CUDA tensor."""
    _lib.require_cuda(img_u8)
    return _canny(img_u8.contiguous(), low, high)


@_lib.on_tensor_device
def sobel3(img: torch.Tensor):
    """(cv2.Sobel(img, CV_32F, 1, 0, ksize=3), cv2.Sobel(img, CV_32F, 0, 1, ksize=3)) for a float32 [H,W] CUDA tensor."""
    _lib.require_cuda(img)
    x = img.float().contiguous()
    H, W = x.shape
    dx, dy = torch.empty_like(x), torch.empty_like(x)
    _lib.check(_lib.lib().t3d_sobel3_f32(_lib.ptr(x), _lib.ptr(dx), _lib.ptr(dy), H, W, _lib.current_stream_ptr()), "t3d_sobel3_f32")
    return dx, dy


@_lib.on_tensor_device
def bilateral_filter(img: torch.Tensor, d: int, sigma_color: float, sigma_space: float) -> torch.Tensor:
    """cv2.bilateralFilter for float32 [H,W] or [H,W,3] CUDA tensors."""
    _lib.require_cuda(img)
    return _bilateral(img.float().contiguous(), d, sigma_color, sigma_space)


@_lib.on_tensor_device
def histogram100(x: torch.Tensor) -> torch.Tensor:
    """np.histogram(x, bins=100, range=(0, 1))[0] as an int32 CUDA tensor."""
    _lib.require_cuda(x)
    v = x.float().contiguous()
    hist = torch.empty(100, dtype=torch.int32, device=v.device)
    _lib.check(_lib.lib().t3d_histogram100(_lib.ptr(v), v.numel(), _lib.ptr(hist), _lib.current_stream_ptr()), "t3d_histogram100")
    return hist
