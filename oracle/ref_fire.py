"""Oracle (test infrastructure) for the experimental fire-scene pipeline,
/root/reference/thermal_dustr_inference_for_experiment.py:62-377.

The reference builds these functions from third-party calls (OpenCV 4.10 pinned by the reference, 4.13 in this image;
NumPy; SciPy): `cv2.createCLAHE`, `cv2.Canny`, `cv2.Sobel`, `cv2.bilateralFilter`, `np.percentile`, `np.histogram`,
`scipy.signal.find_peaks`.  Two layers here:

* plain-NumPy restatements of the libraries' published algorithms (`clahe_u8`, `canny_u8`, `sobel3`, `histogram100`,
  `find_peaks_height_distance`, `bilateral`, `outlier_median`) -- what the CUDA kernels implement; pinned against the
  live libraries in tests/test_oracle_pin.py (CLAHE / Canny / histogram / peaks: identical; Sobel / bilateral: to rounding);
* the three reference functions restated line by line on top of the stock libraries (`preprocess_fire_scene_thermal`,
  `advanced_fire_scene_processing`, `depth_refinement`), with the `np.random.rand` texture passed in (the drop-in draws
  it from the same global NumPy generator at the same point, so a seeded run is reproducible on both sides) and
  without `cv2.ximgproc.guidedFilter` (absent from this image: no oracle).
"""
import numpy as np

try:
    import cv2
    cv2.ipp.setUseIPP(False)
except Exception:           # pragma: no cover
    cv2 = None


# ----------------------------------------------------------------------------- library algorithms, restated
def clahe_u8(src, clip, tiles=(8, 8)):
    """cv2.createCLAHE(clipLimit=clip, tileGridSize=tiles).apply(src) (modules/imgproc/src/clahe.cpp)."""
    tx, ty = tiles
    H, W = src.shape
    ext = src
    if W % tx or H % ty:                                       # copyMakeBorder(..., BORDER_REFLECT_101) at the bottom / right
        ext = np.pad(src, ((0, (ty - H % ty) % ty), (0, (tx - W % tx) % tx)), mode="reflect")
    th, tw = ext.shape[0] // ty, ext.shape[1] // tx
    area = th * tw
    lut_scale = np.float32(255.0) / np.float32(area)
    cl = max(int(clip * area / 256), 1) if clip > 0 else 0
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            h = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if cl > 0:
                clipped = int(np.maximum(h - cl, 0).sum())
                h = np.minimum(h, cl)
                batch = clipped // 256
                residual = clipped - batch * 256
                h += batch
                if residual:
                    step = max(256 // residual, 1)
                    k = 0
                    while k < 256 and residual > 0:
                        h[k] += 1; k += step; residual -= 1
            luts[j, i] = np.clip(np.rint(np.cumsum(h).astype(np.float32) * lut_scale), 0, 255).astype(np.uint8)
    f32 = np.float32
    xs = np.arange(W, dtype=f32) * (f32(1) / f32(tw)) - f32(0.5)
    tx1 = np.floor(xs).astype(np.int32); xa = (xs - tx1.astype(f32)).astype(f32); xa1 = (f32(1) - xa).astype(f32)
    tx2 = np.minimum(tx1 + 1, tx - 1); tx1 = np.maximum(tx1, 0)
    ys = np.arange(H, dtype=f32) * (f32(1) / f32(th)) - f32(0.5)
    ty1 = np.floor(ys).astype(np.int32); ya = (ys - ty1.astype(f32)).astype(f32); ya1 = (f32(1) - ya).astype(f32)
    ty2 = np.minimum(ty1 + 1, ty - 1); ty1 = np.maximum(ty1, 0)
    out = np.zeros_like(src)
    for y in range(H):
        v = src[y].astype(np.int64)
        l11, l12 = luts[ty1[y], tx1, v].astype(f32), luts[ty1[y], tx2, v].astype(f32)
        l21, l22 = luts[ty2[y], tx1, v].astype(f32), luts[ty2[y], tx2, v].astype(f32)
        res = (l11 * xa1 + l12 * xa) * ya1[y] + (l21 * xa1 + l22 * xa) * ya[y]
        out[y] = np.clip(np.rint(res), 0, 255).astype(np.uint8)
    return out


def canny_u8(img, low, high):
    """cv2.Canny(img, low, high): aperture 3, L1 gradient (modules/imgproc/src/canny.cpp)."""
    low, high = int(np.floor(low)), int(np.floor(high))
    if low > high:
        low, high = high, low
    p = np.pad(img.astype(np.int32), 1, mode="edge")             # Sobel with BORDER_REPLICATE
    rx = p[:, 2:] - p[:, :-2]
    dx = rx[:-2] + 2 * rx[1:-1] + rx[2:]
    ry = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    dy = ry[2:] - ry[:-2]
    mag = np.abs(dx) + np.abs(dy)
    mp = np.pad(mag, 1)                                           # zero border
    m = mp[1:-1, 1:-1]
    x = np.abs(dx).astype(np.int64); y = np.abs(dy).astype(np.int64) << 15
    tg22 = x * 13573; tg67 = tg22 + (x << 16)
    left, right, up, down = mp[1:-1, :-2], mp[1:-1, 2:], mp[:-2, 1:-1], mp[2:, 1:-1]
    neg = (dx ^ dy) < 0                                           # s = -1: prev[j + 1], next[j - 1]
    prev_d = np.where(neg, mp[:-2, 2:], mp[:-2, :-2]); next_d = np.where(neg, mp[2:, :-2], mp[2:, 2:])
    horiz = y < tg22; vert = (~horiz) & (y > tg67); diag = (~horiz) & (~vert)
    keep = (horiz & (m > left) & (m >= right)) | (vert & (m > up) & (m >= down)) | (diag & (m > prev_d) & (m > next_d))
    cand = (m > low) & keep
    lab = cand & (m > high)
    while True:                                                   # hysteresis: candidates 8-connected to a seed
        g = np.pad(lab, 1)
        nb = g[:-2, :-2] | g[:-2, 1:-1] | g[:-2, 2:] | g[1:-1, :-2] | g[1:-1, 2:] | g[2:, :-2] | g[2:, 1:-1] | g[2:, 2:]
        new = lab | (cand & nb)
        if (new == lab).all():
            break
        lab = new
    return (lab * 255).astype(np.uint8)


def sobel3(img):
    """cv2.Sobel(img, CV_32F, 1, 0, ksize=3), cv2.Sobel(img, CV_32F, 0, 1, ksize=3) (BORDER_REFLECT_101); agrees with
    OpenCV's separable SIMD evaluation to one rounding."""
    p = np.pad(img.astype(np.float32), 1, mode="reflect")
    two = np.float32(2)
    rx = p[:, 2:] - p[:, :-2]
    dx = rx[1:-1] * two + (rx[:-2] + rx[2:])
    ry = p[:, 1:-1] * two + (p[:, :-2] + p[:, 2:])
    return dx.astype(np.float32), (ry[2:] - ry[:-2]).astype(np.float32)


def histogram100(x):
    """np.histogram(x, bins=100, range=(0, 1))[0] (numpy/lib/_histograms_impl.py, the equal-width fast path)."""
    a = np.asarray(x).astype(np.float64).ravel()
    edges = np.linspace(0.0, 1.0, 101)
    a = a[(a >= 0.0) & (a <= 1.0)]
    idx = ((a - 0.0) / (1.0 - 0.0) * 100).astype(np.intp)
    idx[idx == 100] -= 1
    idx[a < edges[idx]] -= 1
    inc = (a >= edges[idx + 1]) & (idx != 99)
    idx[inc] += 1
    return np.bincount(idx, minlength=100).astype(np.int64)


def find_peaks_height_distance(x, height, distance):
    """scipy.signal.find_peaks(x, height=height, distance=distance)[0] (_local_maxima_1d, height filter,
    _select_by_peak_distance with scipy's own priority order: np.argsort of the heights, highest first)."""
    x = np.asarray(x, np.float64)
    n = len(x)
    mids = []
    i = 1
    while i < n - 1:
        if x[i - 1] < x[i]:
            a = i + 1
            while a < n - 1 and x[a] == x[i]:
                a += 1
            if x[a] < x[i]:
                mids.append((i + a - 1) // 2)
                i = a
        i += 1
    peaks = np.array([q for q in mids if x[q] >= height], np.intp)
    if len(peaks):
        keep = np.ones(len(peaks), bool)
        order = np.argsort(x[peaks])
        for t in range(len(peaks) - 1, -1, -1):
            j = order[t]
            if not keep[j]:
                continue
            k = j - 1
            while k >= 0 and peaks[j] - peaks[k] < distance:
                keep[k] = False; k -= 1
            k = j + 1
            while k < len(peaks) and peaks[k] - peaks[j] < distance:
                keep[k] = False; k += 1
        peaks = peaks[keep]
    return peaks


def fire_threshold_from_histogram(hist):
    """advanced_fire_scene_processing :188-214 reduced to what the rest of the function uses: the lower bound of the
    highest temperature region (the last mask) -- midpoint of the two highest histogram peaks, else 0.7."""
    hist = np.asarray(hist)
    bins = np.linspace(0.0, 1.0, 101)
    peaks = find_peaks_height_distance(hist, hist.max() * 0.3, 10)
    pv = np.sort(bins[peaks])
    if len(pv) >= 2:
        return float((pv[-2] + pv[-1]) / 2)
    return 0.7


def bilateral(img, d, sigma_color, sigma_space):
    """cv2.bilateralFilter on float32 [H,W] / [H,W,3]: direct evaluation (OpenCV tabulates the colour weight; equal to
    rounding)."""
    x = np.asarray(img, np.float32)
    x3 = x[..., None] if x.ndim == 2 else x
    if abs(float(x3.max()) - float(x3.min())) < np.finfo(np.float32).eps:
        return x.copy()
    radius = max(d // 2 if d > 0 else int(round(sigma_space * 1.5)), 1)
    gc, gs = -0.5 / (sigma_color * sigma_color), -0.5 / (sigma_space * sigma_space)
    H, W = x3.shape[:2]
    p = np.pad(x3.astype(np.float64), ((radius, radius), (radius, radius), (0, 0)), mode="reflect")
    num = x3.astype(np.float64).copy()
    den = np.ones((H, W, 1))
    for dy in range(-radius, radius + 1):
        for dx in range(-radius, radius + 1):
            r2 = dy * dy + dx * dx
            if r2 == 0 or r2 > radius * radius:
                continue
            v = p[radius + dy:radius + dy + H, radius + dx:radius + dx + W]
            w = np.exp(r2 * gs) * np.exp(np.abs(v - x3).sum(-1, keepdims=True) ** 2 * gc)
            num += v * w
            den += w
    out = (num / den).astype(np.float32)
    return out[..., 0] if x.ndim == 2 else out


def outlier_median(depth):
    """depth_refinement_with_outlier_removal :335-356 (the explicit double loop, vectorised over the outliers only)."""
    depth = np.asarray(depth)
    mean, std = np.nanmean(depth), np.nanstd(depth)
    mask = np.abs(depth - mean) > 3 * std
    out = np.copy(depth)
    H, W = depth.shape
    for i, j in zip(*np.nonzero(mask)):
        nb = depth[max(0, i - 2):min(H, i + 3), max(0, j - 2):min(W, j + 3)]
        nb = nb[~mask[max(0, i - 2):min(H, i + 3), max(0, j - 2):min(W, j + 3)]]
        out[i, j] = np.median(nb) if nb.size > 0 else mean
    return out, mask, mean, std


# ----------------------------------------------------------------------------- the reference functions on the stock libraries
def _gray(thermal_chw):
    t = np.asarray(thermal_chw, np.float32)
    if t.ndim == 3 and t.shape[0] == 3:
        t = t.transpose(1, 2, 0)
    if t.ndim == 3 and t.shape[2] >= 3:
        return 0.299 * t[:, :, 0] + 0.587 * t[:, :, 1] + 0.114 * t[:, :, 2]                # :86-87
    return t[:, :, 0] if t.ndim == 3 else t


def preprocess_fire_scene_thermal(thermal_chw, noise01):
    """:62-152 with `np.random.rand(h, w).astype(np.float32)` given as `noise01` (None = no texture)."""
    g = _gray(thermal_chw)
    p_low, p_high = np.percentile(g, (5, 95))                                               # :95
    tn = np.clip(g, p_low, p_high)
    tn = (tn - p_low) / (p_high - p_low + 1e-6)
    fire = tn > 0.7
    h, w = tn.shape
    out = np.zeros((h, w, 3), np.float32)
    base = np.clip((1.0 - tn) * 1.2, 0, 1)
    base_clahe = cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply((base * 255).astype(np.uint8)).astype(np.float32) / 255.0
    for c in range(3):
        out[:, :, c] = base_clahe
    out[fire, 0] = 0.8; out[fire, 1] = 0.3; out[fire, 2] = 0.1
    if noise01 is not None:
        noise = noise01.astype(np.float32) * 0.1
        for c in range(3):
            out[:, :, c] = np.where(fire, out[:, :, c] + noise, out[:, :, c])
    edges = cv2.Canny((tn * 255).astype(np.uint8), 50, 150).astype(np.float32) / 255.0
    ew = np.ones_like(tn) * 0.15
    ew[fire] = 0.3
    for c in range(3):
        out[:, :, c] = out[:, :, c] * (1 - ew) + edges * ew
    return np.clip(out, 0, 1).transpose(2, 0, 1)


def advanced_fire_scene_processing(thermal_chw, noise01):
    """:154-282 (scipy.signal.find_peaks through the restatement above)."""
    g = _gray(thermal_chw)
    hist, bins = np.histogram(g.flatten(), bins=100, range=(0, 1))
    thr = fire_threshold_from_histogram(hist)
    fire = g > (np.float64(thr) if thr != 0.7 else 0.7)
    h, w = g.shape
    out = np.zeros((h, w, 3), np.float32)
    inv = 1.0 - g
    cl = cv2.createCLAHE(clipLimit=2.5, tileGridSize=(8, 8)).apply((inv * 255).astype(np.uint8)).astype(np.float32) / 255.0
    e1 = cv2.Canny((g * 255).astype(np.uint8), 30, 150).astype(np.float32) / 255.0
    sx = cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=3); sy = cv2.Sobel(g, cv2.CV_32F, 0, 1, ksize=3)
    sm = np.sqrt(sx ** 2 + sy ** 2)
    sm = (sm - sm.min()) / (sm.max() - sm.min() + 1e-6)
    edges = np.maximum(e1, sm)
    for c in range(3):
        out[:, :, c] = cl
    out[fire, 0] = cl[fire] * 0.5; out[fire, 1] = cl[fire] * 0.3; out[fire, 2] = cl[fire] * 0.2
    if noise01 is not None:
        noise = noise01.astype(np.float32) * 0.15
        for c in range(3):
            out[:, :, c] = np.where(fire, out[:, :, c] + noise, out[:, :, c])
    es = np.ones_like(g) * 0.2
    es[fire] = 0.4
    for c in range(3):
        out[:, :, c] = out[:, :, c] * (1 - es) + edges * es
    out = cv2.bilateralFilter(out, 9, 75, 75)
    return np.clip(out, 0, 1).transpose(2, 0, 1)


def depth_refinement(depth):
    """:284-377 with guided_filter=False: outlier removal + cv2.bilateralFilter(., 5, 50, 50)."""
    cleaned, _, _, _ = outlier_median(np.asarray(depth, np.float32))
    return cv2.bilateralFilter(cleaned.astype(np.float32), 5, 50, 50)


def make_fire_frame(h, w, seed=0):
    """A synthetic [3,H,W] thermal frame in [0, 1]: smooth background, a hot blob, sensor noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    bg = 0.25 + 0.15 * np.sin(xx / (7.0 + seed)) * np.cos(yy / 11.0) + 0.1 * (yy / h)
    blob = 0.6 * np.exp(-(((xx - 0.6 * w) / (0.12 * w)) ** 2 + ((yy - 0.4 * h) / (0.18 * h)) ** 2))
    box = np.zeros((h, w), np.float32)
    box[h // 8: h // 3, w // 10: w // 3] = 0.3                                  # a warm object with sharp borders
    box[(yy + 2 * xx).astype(np.int64) % 37 == 0] += 0.2                        # thin hot streaks
    g = np.clip(bg + blob + box + 0.02 * rng.standard_normal((h, w)), 0, 1).astype(np.float32)
    return np.repeat(g[None], 3, 0).copy()
