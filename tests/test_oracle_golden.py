"""CPU: the oracle restatements reproduce the golden vectors generated from the
UNMODIFIED reference (oracle/gen_golden.py, run in the build container)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_depth, ref_loss, ref_metrics, ref_preprocess, ref_sobel

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KW = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4)
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def loss_kat():
    return np.load(os.path.join(G, "loss_kat.npz"))


def test_loss_table_survey_appendix_c(loss_kat):
    """KAT-L of SURVEY.md Appendix C (the table rows were produced by utils/loss.py itself)."""
    for row in loss_kat["table"]:
        H, W, multi = int(row[0]), int(row[1]), bool(row[2])
        if H * W > 60000:
            continue        # the 384x512 rows are checked on the GPU box and in test_oracle_pin
        a = [x.clone() for x in ref_loss.make_kat_inputs(H, W, seed=0)]
        for k in (0, 1, 4, 5):
            a[k].requires_grad_()
        loss, comp = ref_loss.enhanced_thermal_aware_loss_torch(*a, multi_scale=multi, **KW)
        loss.backward()
        got = [loss.item(), comp["basic_loss"], comp["edge_loss"], comp["smoothness_loss"], comp["detail_loss"],
               a[0].grad.abs().double().sum().item(), a[1].grad.abs().double().sum().item(),
               a[4].grad.abs().double().sum().item(), a[5].grad.abs().double().sum().item()]
        np.testing.assert_allclose(got, row[3:], rtol=1e-5)   # fp32 sum order varies with the torch thread count
    # the numbers quoted in SURVEY.md Appendix C
    r = loss_kat["table"][0]
    assert r[3] == pytest.approx(4.0638237, rel=1e-6) and r[5] == pytest.approx(7.79507637, rel=1e-6)


@pytest.mark.parametrize("tag", ["s38x52_m0", "s38x52_m1", "s37x51_m0", "s37x51_m1", "s16x128_m1", "s9x6_m0", "s9x6_m1"])
def test_loss_small_cases_full_gradients(loss_kat, tag):
    H, W = (int(v) for v in tag[1:].split("_")[0].split("x"))
    multi = tag.endswith("m1")
    seed = {(38, 52): 1, (37, 51): 2, (16, 128): 3, (9, 6): 4}[(H, W)]
    a = [x.clone() for x in ref_loss.make_kat_inputs(H, W, seed=seed)]
    a[4] = a[4] * 3 - 1.0
    for k in (0, 1, 4, 5):
        a[k].requires_grad_()
    loss, comp = ref_loss.enhanced_thermal_aware_loss_torch(*a, multi_scale=multi, **KW)
    loss.backward()
    sc = loss_kat[tag + "_scalars"]
    np.testing.assert_allclose([loss.item(), comp["basic_loss"], comp["edge_loss"], comp["smoothness_loss"],
                                comp["detail_loss"]], sc, rtol=1e-5)
    for k, name in ((0, "dp1"), (1, "dp2"), (4, "dc1"), (5, "dc2")):
        np.testing.assert_allclose(a[k].grad.numpy(), loss_kat[f"{tag}_{name}"], rtol=1e-5, atol=1e-9)
    # the fp64 closed form (forward + hand-derived backward) agrees with the reference's autograd
    b = ref_loss.make_kat_inputs(H, W, seed=seed)
    f = ref_loss.loss_fwd_bwd_f64(b[0], b[1], b[2], b[3], b[4] * 3 - 1.0, b[5], b[6], b[7], multi_scale=multi, **KW)
    assert f["total"] == pytest.approx(sc[0], rel=2e-6)
    for name in ("dp1", "dp2", "dc1", "dc2"):
        ref = loss_kat[f"{tag}_{name}"].astype(np.float64)
        assert np.abs(f[name] - ref).max() <= 2e-4 * np.abs(ref).max()


def test_loss_v1_and_basic(loss_kat):
    a = ref_loss.make_kat_inputs(224, 224, seed=0)
    l, c = ref_loss.thermal_aware_loss_torch(*a, alpha=0.2, edge_weight=0.5, smoothness_weight=0.3)
    np.testing.assert_allclose([l.item(), c["basic_loss"], c["edge_loss"], c["smoothness_loss"]], loss_kat["v1"], rtol=1e-5)
    a = ref_loss.make_kat_inputs(38, 52, seed=1)
    b = ref_loss.confidence_weighted_regression_loss_torch(a[0], a[1], a[2], a[3], a[4], a[5], alpha=0.2).item()
    bn = ref_loss.confidence_weighted_regression_loss_torch(a[0], a[1], a[2], a[3]).item()
    np.testing.assert_allclose([b, bn], loss_kat["basic_only"], rtol=1e-5)


def test_preprocess_golden():
    k = np.load(os.path.join(G, "preprocess_kat.npz"))
    raw = np.random.default_rng(0).normal(22800, 400, (512, 640)).clip(0, 65535).astype(np.uint16)
    assert sha(raw) == str(k["raw_sha"]) == "a85ab479a0ae4017"          # SURVEY.md Appendix C KAT-P
    mix = ref_preprocess.make_raw_frames(3, seed=7)
    for name, frame in (("day0", raw), ("mix0", mix[0]), ("mix1", mix[1]), ("mix2", mix[2])):
        assert sha(frame) == str(k[name + "_raw_sha"])
        for (w, h) in ((224, 224), (512, 384), (333, 217)):
            tag = f"{name}_{w}x{h}"
            out, p2, p98, r16 = ref_preprocess.train_path(frame, (h, w))
            assert [sha(r16), sha(out)] == list(k[tag + "_train"])
            assert (p2, p98) == tuple(k[tag + "_train_p"])
            assert sha(ref_preprocess.histogram_u16(r16)) == str(k[tag + "_hist_sha"])
            outi, q2, q98, rf = ref_preprocess.inference_path(frame, (h, w))
            assert [sha(rf), sha(outi)] == list(k[tag + "_infer"])
            assert (q2, q98) == tuple(k[tag + "_infer_p"])
            ti = np.repeat(rf[None], 3, 0)
            tt = np.repeat(r16.astype(np.float32)[None], 3, 0)
            assert [sha(ref_preprocess.enhance_thermal_fixed_range(ti)),
                    sha(ref_preprocess.enhance_thermal_fixed_range(tt, normalized=False))] == list(k[tag + "_fixed"])
    assert (ref_preprocess.resize_bilinear(k["small_raw"], (28, 36)) == k["small_resized_u16"]).all()
    assert (ref_preprocess.train_path(k["small_raw"], (28, 36))[0] == k["small_train_out"]).all()
    assert (ref_preprocess.inference_path(k["small_raw"], (28, 36))[0] == k["small_infer_out"]).all()
    assert (ref_preprocess.enhance_thermal_contrast(k["gray_in"])[0] == k["gray_out"]).all()
    assert (ref_preprocess.enhance_thermal_fixed_range(k["gray_in"]) == k["gray_fixed_out"]).all()
    assert (ref_preprocess.resize_nearest(k["nearest_in"], (21, 33)) == k["nearest_out"]).all()


def test_metrics_golden():
    k = np.load(os.path.join(G, "metrics_kat.npz"))
    # KAT-M1 of SURVEY.md Appendix C (full 512x512 fixture pair, values produced by utils/metrics.py)
    np.testing.assert_allclose(k["day_full_ms1"][:4], [0.26554698, 1.3391004, 4.2078977, 0.39255357], rtol=1e-6)
    assert k["day_full_ms1"][4] == 0.5627975463867188
    for split in ("day", "night"):
        gt, pred = k[f"{split}_gt"], k[f"{split}_pred"]
        for ms in (True, False):
            m = ref_metrics.compute_depth_metrics(pred, gt, median_scaling=ms)
            np.testing.assert_array_equal([m[x] for x in ref_metrics.KEYS7], k[f"{split}_crop_ms{int(ms)}"])
            e = ref_metrics.compute_depth_metrics_eval(pred, gt, median_scaling=ms)
            np.testing.assert_array_equal([e["rmse"], e["acc_1.25"], e["acc_1.25^2"]], k[f"{split}_crop_eval_ms{int(ms)}"])
        m = ref_metrics.compute_depth_metrics(pred, gt, mask=k[f"{split}_mask"])
        np.testing.assert_array_equal([m[x] for x in ref_metrics.KEYS7], k[f"{split}_crop_masked"])
        m = ref_metrics.compute_depth_metrics(pred, ref_depth.match_gt(k[f"{split}_gt_big"], pred.shape))
        np.testing.assert_array_equal([m[x] for x in ref_metrics.KEYS7], k[f"{split}_crop_resampled"])
    empty = ref_metrics.compute_depth_metrics(k["day_pred"], np.zeros_like(k["day_gt"]))
    assert set(empty) == {"abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3"} and np.isnan(empty["rmse"])
    # calibration KATs
    assert k["calib_json_K"][0, 0] == 465.2095642089844 and k["calib_json_K"][1, 2] == 249.76126098632812
    assert k["calib_yaml_Kl"][0, 0] == 510.09593415566053
    np.testing.assert_array_equal(ref_depth.estimate_camera_intrinsics(k["focal_pointmap"], k["focal_pointmap"][..., 2]), k["focal_K"])


def test_metric_accumulator_semantics():
    ms = [{k: 1.0 for k in ref_metrics.KEYS7}, {k: 3.0 for k in ref_metrics.KEYS7}]
    ms[1]["rmse"] = np.nan                 # non-finite skipped but still counted (utils/metrics.py:129-136)
    avg = ref_metrics.accumulate_dataset(ms)
    assert avg["abs_rel"] == 2.0 and avg["rmse"] == 0.5


def test_sobel_golden():
    k = np.load(os.path.join(G, "sobel_kat.npz"))
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    y = ref_sobel.preprocess_thermal_torch(x, torch.tensor(0.5), torch.tensor(1.0))
    assert sha(y.numpy()) == str(k["big_sha"])
    assert y.double().sum().item() == pytest.approx(267724.249, rel=1e-8)
    xs = torch.from_numpy(k["small_in"]).requires_grad_()
    ew, ts = torch.tensor(0.8, requires_grad=True), torch.tensor(0.9, requires_grad=True)
    ys = ref_sobel.preprocess_thermal_torch(xs, ew, ts)
    (ys * torch.from_numpy(k["small_w"])).sum().backward()
    np.testing.assert_array_equal(ys.detach().numpy(), k["small_out"])
    np.testing.assert_allclose(xs.grad.numpy(), k["small_dx"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose([ew.grad.item(), ts.grad.item()], [k["small_dew"], k["small_dts"]], rtol=1e-6)
