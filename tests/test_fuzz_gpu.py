"""GPU: randomised shapes through the fast paths against the general paths / the oracle.  Every case is tiny; the
point is odd sizes, partial strips and tiles, band boundaries, resize ratios on both sides of 1 and 2."""
import numpy as np
import pytest
import torch

from oracle import ref_loss, ref_metrics, ref_preprocess

pytestmark = pytest.mark.gpu
KW = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4)


def test_preprocess_random_geometries(cuda_device):
    """Sampled-window percentiles + marching resize vs the exact-histogram path (bit for bit) and the oracle."""
    from thermal3d_vision_b200 import preprocessing as pp
    rng = np.random.default_rng(101)
    for case in range(14):
        sh, sw = int(rng.integers(24, 200)), int(rng.integers(3, 40)) * 8          # fast path needs sw % 8 == 0
        dh, dw = int(rng.integers(9, 260)), int(rng.integers(3, 90)) * 4           # ... and dw % 4 == 0
        if case % 5 == 4:
            sw += 3; dw += 1                                                       # general (scalar) path too
        B = int(rng.integers(1, 4))
        kind = case % 3
        if kind == 0:
            raw = rng.normal(22800, 400, (B, sh, sw)).clip(0, 65535).astype(np.uint16)
        elif kind == 1:
            raw = rng.integers(0, 65536, (B, sh, sw)).astype(np.uint16)
        else:
            raw = (20000 + 3000 * np.sin(np.arange(sw)[None, None, :] / 7.0) * np.cos(np.arange(sh)[None, :, None] / 5.0)
                   + rng.normal(0, 30, (B, sh, sw))).clip(0, 65535).astype(np.uint16)
        d = torch.from_numpy(raw).to(cuda_device)
        a = pp.preprocess_thermal_batch(d, (dw, dh), path="train", histogram=True)
        b = pp.preprocess_thermal_batch(d, (dw, dh), path="train", histogram=False)
        tag = (case, sh, sw, dh, dw)
        assert torch.equal(a.percentiles, b.percentiles), tag
        assert torch.equal(a.thermal.view(torch.int32), b.thermal.view(torch.int32)), tag
        o, p2, p98, _ = ref_preprocess.train_path(raw[0], (dh, dw))
        assert np.array_equal(b.thermal[0].cpu().numpy(), o, equal_nan=True), tag
        if a.grad_stats is not None:
            assert torch.equal(a.grad_stats, b.grad_stats), tag


def test_metrics_random_shapes_and_resampled_gt(cuda_device):
    """Fast extraction (AoS / planar, same-size and nearest-resampled GT) vs the oracle, incl. invalid regions."""
    from thermal3d_vision_b200 import metrics as tm
    rng = np.random.default_rng(202)
    for case in range(12):
        H, W = int(rng.integers(8, 120)), int(rng.integers(2, 40)) * 4
        gh, gw = (H, W) if case % 2 == 0 else (int(rng.integers(8, 150)), int(rng.integers(8, 150)))
        B = int(rng.integers(1, 4))
        gt = (1.5 + 3 * np.abs(rng.standard_normal((B, gh, gw)))).astype(np.float32)
        gt[rng.random((B, gh, gw)) < 0.1] = 0.0
        gts = np.stack([ref_preprocess.resize_nearest(g, (H, W)) for g in gt]) if (gh, gw) != (H, W) else gt
        pm = rng.standard_normal((B, H, W, 3)).astype(np.float32)
        pm[..., 2] = np.where(gts > 0, gts, 1.0) * (1 + 0.1 * rng.standard_normal((B, H, W))).astype(np.float32) * 1.7
        pred = torch.from_numpy(pm).to(cuda_device) if case % 3 else torch.from_numpy(pm[..., 2].copy()).to(cuda_device)
        got = tm.compute_depth_metrics_batch(pred, torch.from_numpy(gt).to(cuda_device))["metrics_f64"].cpu().numpy()
        for b in range(B):
            if not (gts[b] > 0).any():
                continue
            ref = ref_metrics.compute_depth_metrics(pm[b, ..., 2], gts[b])
            np.testing.assert_allclose(got[b, :4], [ref[k] for k in ref_metrics.KEYS7[:4]], rtol=1e-5, err_msg=str((case, H, W, gh, gw)))
            np.testing.assert_array_equal(got[b, 4:7], [ref[k] for k in ref_metrics.KEYS7[4:]], err_msg=str((case, H, W, gh, gw)))


def test_loss_random_shapes_both_scales(cuda_device):
    from thermal3d_vision_b200 import loss as t3d
    rng = np.random.default_rng(303)
    for case in range(10):
        H, W = int(rng.integers(4, 90)), int(rng.integers(1, 50)) * 4
        multi = bool(case % 2)
        B = int(rng.integers(1, 4))
        P1, P2, G1, G2, C1, C2, T1, T2 = ref_loss.make_batch_inputs(B, H, W, seed=1000 + case, stress_conf=True)
        Pa, Pb, Ca, Cb = (x.clone().requires_grad_() for x in (P1, P2, C1, C2))
        mean, rows, valid = ref_loss.batched_loss_torch(Pa, Pb, G1, G2, Ca, Cb, T1, T2, multi_scale=multi, **KW)
        mean.backward()
        d = [x.to(cuda_device) for x in (P1, P2, G1, G2, C1, C2, T1, T2)]
        for k in (0, 1, 4, 5):
            d[k].requires_grad_()
        res = t3d.fused_thermal_loss(*d, multi_scale=multi, **KW)
        res.loss.backward()
        tag = str((case, B, H, W, multi))
        np.testing.assert_allclose(res.per_sample[:, :5].cpu().numpy(), rows, rtol=1e-5, err_msg=tag)
        for got, ref in ((d[0].grad, Pa.grad), (d[1].grad, Pb.grad), (d[4].grad, Ca.grad), (d[5].grad, Cb.grad)):
            torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-6, msg=lambda m: f"{tag}: {m}")
