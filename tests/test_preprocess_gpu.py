"""GPU parity: thermal preprocessing through the C ABI vs the oracle / golden vectors.
Bar: BIT-EXACT outputs, percentiles and histograms (BASELINE.json north_star)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_preprocess

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(G, "preprocess_kat.npz"))


def _frames():
    raw = np.random.default_rng(0).normal(22800, 400, (512, 640)).clip(0, 65535).astype(np.uint16)
    mix = ref_preprocess.make_raw_frames(3, seed=7)
    return [("day0", raw), ("mix0", mix[0]), ("mix1", mix[1]), ("mix2", mix[2])]


@pytest.mark.parametrize("w,h", [(224, 224), (512, 384), (333, 217)])
def test_train_path_bit_exact_vs_reference_golden(cuda_device, kat, w, h):
    """data/dataset_loader.py:237-249 + enhance_thermal_contrast, batched over 4 frames."""
    from thermal3d_vision_b200 import preprocessing as pp
    frames = _frames()
    raw = torch.from_numpy(np.stack([f for _, f in frames])).to(cuda_device)
    tb = pp.preprocess_thermal_batch(raw, (w, h), path="train")
    r16 = pp.resize_bilinear(raw, (h, w), mode="u16")
    out = tb.thermal.cpu().numpy()
    hist = tb.histogram.cpu().numpy().view(np.uint32)
    pct = tb.percentiles.cpu().numpy()
    for i, (name, frame) in enumerate(frames):
        tag = f"{name}_{w}x{h}"
        assert [sha(r16[i].cpu().numpy()), sha(out[i])] == list(kat[tag + "_train"]), tag
        assert tuple(pct[i]) == tuple(kat[tag + "_train_p"]), tag
        assert sha(hist[i]) == str(kat[tag + "_hist_sha"]), tag
        assert float(out[i].astype(np.float64).sum()) == float(kat[tag + "_train_sum"])
        # and against the oracle arrays directly
        o, p2, p98, r = ref_preprocess.train_path(frame, (h, w))
        assert (out[i] == o).all() and (hist[i] == ref_preprocess.histogram_u16(r)).all()
        assert (out[i][0] == out[i][1]).all() and (out[i][0] == out[i][2]).all()


@pytest.mark.parametrize("w,h", [(224, 224), (512, 384), (333, 217)])
def test_inference_path_bit_exact(cuda_device, kat, w, h):
    """thermal_dustr_inference.py:25-60: /65535 -> float resize -> enhance (radix-select percentiles)."""
    from thermal3d_vision_b200 import preprocessing as pp
    frames = _frames()
    raw = torch.from_numpy(np.stack([f for _, f in frames])).to(cuda_device)
    tb = pp.preprocess_thermal_batch(raw, (w, h), path="inference")
    rf = pp.resize_bilinear(raw, (h, w), mode="u16_to_unit_f32")
    out = tb.thermal.cpu().numpy()
    pct = tb.percentiles.cpu().numpy()
    for i, (name, _) in enumerate(frames):
        tag = f"{name}_{w}x{h}"
        assert [sha(rf[i].cpu().numpy()), sha(out[i])] == list(kat[tag + "_infer"]), tag
        assert tuple(pct[i]) == tuple(kat[tag + "_infer_p"]), tag


def test_drop_in_functions(cuda_device, kat):
    from thermal3d_vision_b200 import preprocessing as pp
    # enhance_thermal_contrast: CUDA in -> CUDA out, CPU in -> CPU out, None -> None
    t = torch.from_numpy(np.repeat(kat["small_resized_u16"].astype(np.float32)[None], 3, 0))
    y = pp.enhance_thermal_contrast(t)
    assert not y.is_cuda and y.shape == (3, 28, 36) and y.dtype == torch.float32
    assert (y.numpy() == kat["small_train_out"]).all()
    yc = pp.enhance_thermal_contrast(t.to(cuda_device))
    assert yc.is_cuda and (yc.cpu().numpy() == kat["small_train_out"]).all()
    ti = torch.from_numpy(np.repeat(kat["small_resized_f32"][None], 3, 0))
    assert (pp.enhance_thermal_contrast(ti).numpy() == kat["small_infer_out"]).all()
    assert pp.enhance_thermal_contrast(None) is None and pp.enhance_thermal_fixed_range(None) is None
    # gray path (channels not allclose)
    g = torch.from_numpy(kat["gray_in"])
    assert (pp.enhance_thermal_contrast(g).numpy() == kat["gray_out"]).all()
    assert (pp.enhance_thermal_fixed_range(g).numpy() == kat["gray_fixed_out"]).all()
    # 1-channel and 2-D inputs follow the reference's shape rules
    one = ti[:1]
    o1, _, _ = ref_preprocess.enhance_thermal_contrast(one.numpy())
    y1 = pp.enhance_thermal_contrast(one)
    assert y1.shape == (1, 28, 36) and (y1.numpy() == o1).all()
    two = ti[0]
    o2, _, _ = ref_preprocess.enhance_thermal_contrast(two.numpy())
    y2 = pp.enhance_thermal_contrast(two)
    assert y2.shape == (3, 28, 36) and (y2.numpy() == o2).all()
    # fixed range, both modes and shapes
    for x, norm in ((ti, True), (t, False), (one, True), (two, True)):
        ref = ref_preprocess.enhance_thermal_fixed_range(x.numpy(), normalized=norm)
        got = pp.enhance_thermal_fixed_range(x, normalized=norm)
        assert got.shape == ref.shape and (got.numpy() == ref).all()
    # nearest + plain resize helpers
    assert (pp.resize_nearest(torch.from_numpy(kat["nearest_in"]), (21, 33)).numpy() == kat["nearest_out"]).all()
    assert (pp.resize_bilinear(torch.from_numpy(kat["small_raw"]), (28, 36)).numpy() == kat["small_resized_u16"]).all()


def test_edge_cases(cuda_device):
    from thermal3d_vision_b200 import preprocessing as pp
    # constant image: p98 == p2 -> 0/0 = NaN everywhere (utils/preprocessing.py:23)
    c = torch.full((3, 8, 12), 7.0)
    y = pp.enhance_thermal_contrast(c)
    assert torch.isnan(y).all()
    # NaN in the input -> NaN percentiles -> all NaN
    n = torch.rand(3, 8, 12); n[:, 2, 3] = float("nan")
    assert torch.isnan(pp.enhance_thermal_contrast(n)).all()
    # two-valued image -> +-inf / clip behaviour equals numpy
    rng = np.random.default_rng(4)
    for shape in ((3, 5, 7), (3, 1, 9), (3, 31, 2)):
        x = np.repeat(rng.random(shape[1:]).astype(np.float32)[None], 3, 0)
        ref, _, _ = ref_preprocess.enhance_thermal_contrast(x)
        got = pp.enhance_thermal_contrast(torch.from_numpy(x)).numpy()
        assert (got == ref).all()
    # saturated / extreme 16-bit frames
    raw = np.zeros((3, 64, 80), np.uint16)
    raw[0] = 65535; raw[1, ::2] = 65535; raw[2] = rng.integers(0, 65536, (64, 80))
    tb = pp.preprocess_thermal_batch(torch.from_numpy(raw), (40, 32), path="train")
    for i in range(3):
        o, p2, p98, r = ref_preprocess.train_path(raw[i], (32, 40))
        got = tb.thermal[i].cpu().numpy()
        assert np.array_equal(got, o, equal_nan=True)
        assert (tb.histogram[i].cpu().numpy().view(np.uint32) == ref_preprocess.histogram_u16(r)).all()
    # no resize (src size == dst size)
    tb = pp.preprocess_thermal_batch(torch.from_numpy(raw[2:]), (80, 64), path="train")
    o, _, _, _ = ref_preprocess.train_path(raw[2], (64, 80))
    assert (tb.thermal[0].cpu().numpy() == o).all()


def test_full_size_properties(cuda_device):
    """BASELINE config 5 shape: 640x512 u16 -> 512x384, a batch of 32 frames; size-independent properties."""
    from thermal3d_vision_b200 import preprocessing as pp
    raw = ref_preprocess.make_raw_frames(32, seed=21)
    d = torch.from_numpy(raw).to(cuda_device)
    a = pp.preprocess_thermal_batch(d, (512, 384), path="train")
    b = pp.preprocess_thermal_batch(d, (512, 384), path="train")
    assert torch.equal(a.thermal, b.thermal) and torch.equal(a.histogram, b.histogram)      # deterministic
    assert (a.histogram.sum(1) == 384 * 512).all()                                           # checksum of checksums
    assert a.thermal.min() >= 0 and a.thermal.max() <= 1
    # idempotence of the normalisation on its own output range: re-normalising keeps order
    frac = ((a.thermal[:, 0] > 0) & (a.thermal[:, 0] < 1)).float().mean(dim=(1, 2))
    assert ((frac > 0.94) & (frac <= 0.9601)).all()          # 2nd..98th percentile window
    for i in (0, 17, 31):
        o, p2, p98, _ = ref_preprocess.train_path(raw[i], (384, 512))
        assert (a.thermal[i].cpu().numpy() == o).all()
        assert tuple(a.percentiles[i].tolist()) == (p2, p98)


def test_grad_stats_feed_the_loss(cuda_device):
    """The normalisation kernel's thermal-gradient sums equal what the loss's own statistics pass
    computes, and a loss call that consumes them matches the oracle (rtol 1e-5 / grads 1e-4)."""
    from oracle import ref_loss
    from thermal3d_vision_b200 import loss as t3d
    from thermal3d_vision_b200 import preprocessing as pp
    B, H, W = 3, 96, 160
    raw = torch.from_numpy(ref_preprocess.make_raw_frames(2 * B, seed=5)).to(cuda_device)
    tb1 = pp.preprocess_thermal_batch(raw[:B], (W, H), path="train")
    tb2 = pp.preprocess_thermal_batch(raw[B:], (W, H), path="train")
    assert tb1.grad_stats is not None and tb1.grad_stats.shape[0] == B
    P1, P2, G1, G2, C1, C2, _, _ = ref_loss.make_batch_inputs(B, H, W, seed=9)
    d = [x.to(cuda_device) for x in (P1, P2, G1, G2, C1, C2)]
    kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=False)
    a = t3d.fused_thermal_loss_fwd_bwd(*d, tb1.thermal, tb2.thermal, **kw)
    b = t3d.fused_thermal_loss_fwd_bwd(*d, tb1.thermal, tb2.thermal, thermal_stats=(tb1.grad_stats, tb2.grad_stats), **kw)
    torch.testing.assert_close(a["per_sample"], b["per_sample"], rtol=2e-6, atol=0)
    torch.testing.assert_close(a["dpred1"], b["dpred1"], rtol=1e-5, atol=1e-10)
    # sums of the stats against numpy on the produced thermal image
    th = tb1.thermal.cpu().numpy()
    g = (np.float32(0.299) * th[:, 0] + np.float32(0.587) * th[:, 1]) + np.float32(0.114) * th[:, 2]
    sx = np.abs(np.diff(g.astype(np.float64), axis=2)).sum(axis=(1, 2))
    sy = np.abs(np.diff(g.astype(np.float64), axis=1)).sum(axis=(1, 2))
    got = tb1.grad_stats.double().sum(1).cpu().numpy()
    np.testing.assert_allclose(got[:, 0], sx, rtol=1e-5)
    np.testing.assert_allclose(got[:, 1], sy, rtol=1e-5)
    # and the whole thing against the oracle
    mean, rows, _ = ref_loss.batched_loss_torch(P1, P2, G1, G2, C1, C2, tb1.thermal.cpu(), tb2.thermal.cpu(),
                                                alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4,
                                                multi_scale=False)
    np.testing.assert_allclose(b["per_sample"][:, :5].cpu().numpy(), rows, rtol=1e-5)


@pytest.mark.parametrize("w,h,hist", [(160, 96, True), (512, 384, False), (224, 224, False), (136, 50, False)])
def test_half_resolution_grad_stats_feed_the_multi_scale_loss(cuda_device, w, h, hist):
    """half_res_stats=True: the normalisation kernel also sums |Dx|, |Dy| of the 2x2 average-pooled gray image
    (utils/loss.py:133-174); they equal numpy on the produced image, and a multi-scale loss call that consumes them
    matches the oracle.  Shapes the half-resolution pass does not take (odd bands) report stats_scales == 1 and the
    loss computes its own statistics."""
    from oracle import ref_loss
    from thermal3d_vision_b200 import loss as t3d
    from thermal3d_vision_b200 import preprocessing as pp
    B = 2
    raw = torch.from_numpy(ref_preprocess.make_raw_frames(2 * B, seed=15)).to(cuda_device)
    tb1 = pp.preprocess_thermal_batch(raw[:B], (w, h), path="train", histogram=hist, half_res_stats=True)
    tb2 = pp.preprocess_thermal_batch(raw[B:], (w, h), path="train", histogram=hist, half_res_stats=True)
    plain = pp.preprocess_thermal_batch(raw[:B], (w, h), path="train", histogram=hist)
    assert plain.stats_scales == 1 and torch.equal(plain.thermal, tb1.thermal)
    assert torch.equal(plain.grad_stats[..., :2], tb1.grad_stats[..., :2])
    assert float(plain.grad_stats[..., 2:].abs().sum()) == 0.0
    rows_per = -(-h // 24)
    assert tb1.stats_scales == (2 if rows_per % 2 == 0 else 1)
    if tb1.stats_scales == 2:
        th = tb1.thermal.cpu().numpy()
        g = (np.float32(0.299) * th[:, 0] + np.float32(0.587) * th[:, 1]) + np.float32(0.114) * th[:, 2]
        h2, w2 = h // 2, w // 2
        q = g[:, :2 * h2, :2 * w2].reshape(B, h2, 2, w2, 2)
        pooled = (((q[:, :, 0, :, 0] + q[:, :, 0, :, 1]) + q[:, :, 1, :, 0]) + q[:, :, 1, :, 1]) * np.float32(0.25)
        sx2 = np.abs(np.diff(pooled.astype(np.float64), axis=2)).sum(axis=(1, 2))
        sy2 = np.abs(np.diff(pooled.astype(np.float64), axis=1)).sum(axis=(1, 2))
        got = tb1.grad_stats.double().sum(1).cpu().numpy()
        np.testing.assert_allclose(got[:, 2], sx2, rtol=1e-5)
        np.testing.assert_allclose(got[:, 3], sy2, rtol=1e-5)
    P1, P2, G1, G2, C1, C2, _, _ = ref_loss.make_batch_inputs(B, h, w, seed=19)
    d = [x.to(cuda_device) for x in (P1, P2, G1, G2, C1, C2)]
    kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=True)
    out = t3d.fused_thermal_loss_fwd_bwd(*d, tb1.thermal, tb2.thermal, thermal_stats=(tb1.grad_stats, tb2.grad_stats),
                                         thermal_stats_scales=min(tb1.stats_scales, tb2.stats_scales),
                                         thermal_replicated=True, **kw)
    Pa, Pb = P1.clone().requires_grad_(), P2.clone().requires_grad_()
    mean, rows, _ = ref_loss.batched_loss_torch(Pa, Pb, G1, G2, C1, C2, tb1.thermal.cpu(), tb2.thermal.cpu(), **kw)
    mean.backward()
    np.testing.assert_allclose(out["per_sample"][:, :5].cpu().numpy(), rows, rtol=1e-5)
    torch.testing.assert_close(out["dpred1"].cpu(), Pa.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out["dpred2"].cpu(), Pb.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("w,h", [(224, 224), (512, 384), (333, 217), (640, 512)])
def test_bracket_percentiles_match_histogram_path(cuda_device, w, h):
    """histogram=False (sampled value windows, t3d_preprocess_bracket.cu) must give the same bits as the
    exact-histogram path and the oracle: outputs, percentiles, thermal-gradient sums."""
    from thermal3d_vision_b200 import preprocessing as pp
    frames = _frames()
    raw = torch.from_numpy(np.stack([f for _, f in frames])).to(cuda_device)
    a = pp.preprocess_thermal_batch(raw, (w, h), path="train", histogram=True)
    b = pp.preprocess_thermal_batch(raw, (w, h), path="train", histogram=False)
    assert b.histogram is None
    assert torch.equal(a.percentiles, b.percentiles)
    assert torch.equal(a.thermal, b.thermal)
    if a.grad_stats is not None:
        assert torch.equal(a.grad_stats, b.grad_stats)
    for i, (name, frame) in enumerate(frames):
        o, p2, p98, _ = ref_preprocess.train_path(frame, (h, w))
        assert tuple(b.percentiles[i].tolist()) == (p2, p98), name
        assert (b.thermal[i].cpu().numpy() == o).all(), name


def test_bracket_percentiles_adversarial_frames(cuda_device):
    """Frames built to defeat the sampled windows (window wider than its cap, heavy duplicates, two-valued,
    constant, column-periodic): the per-frame exact fallback must kick in and give the oracle's bits."""
    from thermal3d_vision_b200 import preprocessing as pp
    rng = np.random.default_rng(12)
    H, W = 96, 160
    frames = [
        rng.integers(0, 65536, (H, W)).astype(np.uint16),                           # uniform over the whole range
        np.full((H, W), 31000, np.uint16),                                          # constant
        np.where(rng.random((H, W)) < 0.5, 100, 60000).astype(np.uint16),           # two-valued
        (np.arange(W)[None, :] % 16 * 4000 + np.zeros((H, 1))).astype(np.uint16),   # column-periodic
        np.where(rng.random((H, W)) < 0.03, 65535, rng.integers(20000, 20040, (H, W))).astype(np.uint16),
        np.sort(rng.integers(0, 65536, H * W)).reshape(H, W).astype(np.uint16),     # sorted ramp
    ]
    raw = torch.from_numpy(np.stack(frames)).to(cuda_device)
    for size in ((W, H), (80, 64), (224, 224)):          # no resize / downscale / upscale
        a = pp.preprocess_thermal_batch(raw, size, path="train", histogram=True)
        b = pp.preprocess_thermal_batch(raw, size, path="train", histogram=False)
        assert torch.equal(a.percentiles, b.percentiles), size
        assert torch.equal(a.thermal.view(torch.int32), b.thermal.view(torch.int32)), size    # NaN-safe bit compare
        for i, f in enumerate(frames):
            o, p2, p98, _ = ref_preprocess.train_path(f, (size[1], size[0]))
            assert np.array_equal(b.thermal[i].cpu().numpy(), o, equal_nan=True), (size, i)


def test_resize_5_to_4_fast_path_adversarial(cuda_device):
    """The dataset geometry (5:4 on both axes, e.g. 640x512 -> 512x384) takes resize54_march_kernel (compile-time
    taps, below / above counts in registers): bits must equal the histogram path and the oracle on frames that
    stress the windows (full-range noise, two-valued, ramps) and on strips / row ranges of odd sizes."""
    from thermal3d_vision_b200 import preprocessing as pp
    rng = np.random.default_rng(21)
    for (H, W), size in (((80, 160), (128, 64)), ((200, 360), (288, 160)), ((512, 640), (512, 384))):
        frames = [
            rng.integers(0, 65536, (H, W)).astype(np.uint16),
            np.where(rng.random((H, W)) < 0.5, 100, 60000).astype(np.uint16),
            np.sort(rng.integers(0, 65536, H * W)).reshape(H, W).astype(np.uint16),
            rng.normal(22800, 300, (H, W)).clip(0, 65535).astype(np.uint16),
            np.where(rng.random((H, W)) < 0.03, 65535, rng.integers(20000, 20040, (H, W))).astype(np.uint16),
        ]
        raw = torch.from_numpy(np.stack(frames)).to(cuda_device)
        a = pp.preprocess_thermal_batch(raw, size, path="train", histogram=True)
        b = pp.preprocess_thermal_batch(raw, size, path="train", histogram=False)
        assert torch.equal(a.percentiles, b.percentiles), size
        assert torch.equal(a.thermal.view(torch.int32), b.thermal.view(torch.int32)), size
        for i, f in enumerate(frames):
            o, p2, p98, _ = ref_preprocess.train_path(f, (size[1], size[0]))
            assert tuple(b.percentiles[i].tolist()) == (p2, p98), (size, i)
            assert np.array_equal(b.thermal[i].cpu().numpy(), o, equal_nan=True), (size, i)


def test_bracket_windows_hold_on_typical_frames(cuda_device):
    """Performance guard: on ordinary frames (day / night + hot blobs, real-data-like smooth fields) the sampled
    windows must contain the percentile ranks -- no frame may need the exact-select fallback."""
    from thermal3d_vision_b200 import preprocessing as pp
    raw = ref_preprocess.make_raw_frames(48, seed=77)
    yy, xx = np.mgrid[0:512, 0:640]
    smooth = (22500 + 600 * np.sin(xx / 97.0) * np.cos(yy / 61.0) + 40 * np.random.default_rng(5).standard_normal((512, 640)))
    raw = np.concatenate([raw, smooth[None].astype(np.uint16), (smooth[None] * 0 + 23000 + (xx // 8 % 2) * 300).astype(np.uint16)])
    d = torch.from_numpy(raw).to(cuda_device)
    for size in ((512, 384), (224, 224)):
        out2 = {"workspace": torch.empty(pp._lib.lib().t3d_preprocess_workspace_bytes(d.shape[0], size[1], size[0]),
                                         dtype=torch.uint8, device=cuda_device)}
        b = pp.preprocess_thermal_batch(d, size, path="train", histogram=False, out=out2)
        assert pp.bracket_fallback_count(out2["workspace"], d.shape[0], size) == 0, size
        h = pp.preprocess_thermal_batch(d, size, path="train", histogram=True)
        assert torch.equal(b.percentiles, h.percentiles) and torch.equal(b.thermal, h.thermal)


def test_ingest_to_model_input(cuda_device, tmp_path):
    """SURVEY 8f row 3 end to end: PNG-16 files -> native decode into pinned memory -> H2D -> GPU preprocessing
    == cv2.imread + cv2.resize + enhance_thermal_contrast per frame (data/dataset_loader.py:110,237-249)."""
    cv2 = pytest.importorskip("cv2")
    from thermal3d_vision_b200 import ingest
    raw = ref_preprocess.make_raw_frames(4, seed=9)
    paths = []
    for i in range(4):
        p = str(tmp_path / f"fl_ir_aligned_{i}.png")
        assert cv2.imwrite(p, raw[i])
        paths.append(p)
    tb = ingest.load_thermal_batch(paths, img_size=(224, 224), device=cuda_device)
    for i in range(4):
        o, p2, p98, _ = ref_preprocess.train_path(cv2.imread(paths[i], cv2.IMREAD_ANYDEPTH), (224, 224))
        assert (tb.thermal[i].cpu().numpy() == o).all()
        assert tuple(tb.percentiles[i].tolist()) == (p2, p98)


def test_float_percentiles_bracket_and_fallback_paths(cuda_device):
    """enhance_thermal_contrast on float data at sizes where the sampled brackets are used, including inputs that
    overflow the candidate buffers (two-valued, constant, heavy duplicates) and take the exact fallback; a real
    3-channel (non-replicated) image goes through the fp32 gray plane.  Bit-exact vs numpy (oracle)."""
    from thermal3d_vision_b200 import preprocessing as pp
    rng = np.random.default_rng(21)
    H, W = 192, 256
    base = rng.random((H, W)).astype(np.float32)
    cases = {
        "uniform": np.repeat(base[None], 3, 0),
        "gray3": rng.random((3, H, W)).astype(np.float32),
        "quantised": np.repeat((rng.integers(21800, 25000, (H, W)) / 65535.0).astype(np.float32)[None], 3, 0),
        "two_valued": np.repeat(np.where(base < 0.5, 0.25, 0.75).astype(np.float32)[None], 3, 0),
        "constant": np.full((3, H, W), 0.5, np.float32),
        "heavy_dups": np.repeat(np.where(base < 0.9, 0.1, base).astype(np.float32)[None], 3, 0),
        "negative_and_large": np.repeat(((base - 0.5) * 1e6).astype(np.float32)[None], 3, 0),
        "one_channel_2d": base,
    }
    for name, x in cases.items():
        ref, p2, p98 = ref_preprocess.enhance_thermal_contrast(x)
        got = pp.enhance_thermal_contrast(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
        assert got.shape == ref.shape, name
        assert np.array_equal(got, ref, equal_nan=True), name
    # batched inference path (u16 -> /65535 -> float resize -> float percentiles) vs the oracle
    raw = ref_preprocess.make_raw_frames(3, seed=5)
    tb = pp.preprocess_thermal_batch(torch.from_numpy(raw).to(cuda_device), (512, 384), path="inference")
    for i in range(3):
        o = ref_preprocess.inference_path(raw[i], (384, 512))
        o = o[0] if isinstance(o, tuple) else o
        assert np.array_equal(tb.thermal[i].cpu().numpy(), o)
