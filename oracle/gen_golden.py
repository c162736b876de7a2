"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

Usage:  python -m oracle.gen_golden        (from the repo root; needs /root/reference)

Every vector below is an output of the reference's own functions
(utils/loss.py, utils/preprocessing.py, utils/metrics.py,
utils/evaluate_depth_metrics.py, thermal_dustr_model.py, scripts/pseudo_gt.py)
or of the libraries it calls (cv2.resize with IPP off, np.percentile), on
inputs that are either regenerated from a seed by the tests or stored here.
Library versions are recorded in each file.
"""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np
import torch

from . import ref_loss, ref_preprocess, reference_bridge

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def versions(ref):
    return np.array([f"torch {torch.__version__}", f"numpy {np.__version__}", f"cv2 {ref.cv2.__version__} (IPP off)"])


def gen_loss(ref):
    kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4)
    rows = []
    for (H, W) in ((224, 224), (384, 512)):
        for multi in (False, True):
            a = [x.clone() for x in ref_loss.make_kat_inputs(H, W, seed=0)]
            for k in (0, 1, 4, 5):
                a[k].requires_grad_()
            loss, comp = ref.loss.enhanced_thermal_aware_loss(*a, multi_scale=multi, **kw)
            loss.backward()
            rows.append([H, W, int(multi), loss.item(), comp["basic_loss"], comp["edge_loss"],
                         comp["smoothness_loss"], comp["detail_loss"],
                         a[0].grad.abs().double().sum().item(), a[1].grad.abs().double().sum().item(),
                         a[4].grad.abs().double().sum().item(), a[5].grad.abs().double().sum().item()])
    # v1 loss
    a = ref_loss.make_kat_inputs(224, 224, seed=0)
    l1, c1 = ref.loss.thermal_aware_loss(*a, alpha=0.2, edge_weight=0.5, smoothness_weight=0.3)
    v1 = [l1.item(), c1["basic_loss"], c1["edge_loss"], c1["smoothness_loss"]]
    # small cases with full gradients (stress confidences: both clamps hit), odd sizes
    small = {}
    for (H, W, seed) in ((38, 52, 1), (37, 51, 2), (16, 128, 3), (9, 6, 4)):
        for multi in (False, True):
            a = [x.clone() for x in ref_loss.make_kat_inputs(H, W, seed=seed)]
            a[4] = a[4] * 3 - 1.0
            for k in (0, 1, 4, 5):
                a[k].requires_grad_()
            loss, comp = ref.loss.enhanced_thermal_aware_loss(*a, multi_scale=multi, **kw)
            loss.backward()
            tag = f"s{H}x{W}_m{int(multi)}"
            small[tag + "_scalars"] = np.array([loss.item(), comp["basic_loss"], comp["edge_loss"],
                                                comp["smoothness_loss"], comp["detail_loss"]])
            small[tag + "_dp1"] = a[0].grad.numpy()
            small[tag + "_dp2"] = a[1].grad.numpy()
            small[tag + "_dc1"] = a[4].grad.numpy()
            small[tag + "_dc2"] = a[5].grad.numpy()
    # basic-only function
    a = ref_loss.make_kat_inputs(38, 52, seed=1)
    b = ref.loss.confidence_weighted_regression_loss(a[0], a[1], a[2], a[3], a[4], a[5], alpha=0.2).item()
    b_none = ref.loss.confidence_weighted_regression_loss(a[0], a[1], a[2], a[3]).item()
    np.savez_compressed(os.path.join(OUT, "loss_kat.npz"), table=np.array(rows, np.float64), v1=np.array(v1),
                        basic_only=np.array([b, b_none]), versions=versions(ref), **small)


def gen_preprocess(ref):
    cv2 = ref.cv2
    raw = np.random.default_rng(0).normal(22800, 400, (512, 640)).clip(0, 65535).astype(np.uint16)
    night = ref_preprocess.make_raw_frames(3, seed=7)
    out = {"raw_sha": np.array(sha(raw)), "versions": versions(ref)}
    for name, frame in (("day0", raw), ("mix0", night[0]), ("mix1", night[1]), ("mix2", night[2])):
        out[name + "_raw_sha"] = np.array(sha(frame))
        for (w, h) in ((224, 224), (512, 384), (333, 217)):
            tag = f"{name}_{w}x{h}"
            r16 = cv2.resize(frame, (w, h))
            t = torch.from_numpy(np.stack([r16.astype(np.float32)] * 3, -1).transpose(2, 0, 1)).float()
            e = ref.preprocessing.enhance_thermal_contrast(t).numpy()
            out[tag + "_train"] = np.array([sha(r16), sha(e)])
            out[tag + "_train_p"] = np.percentile(t[0].numpy(), (2, 98))
            out[tag + "_train_sum"] = np.array(e.astype(np.float64).sum())
            out[tag + "_hist_sha"] = np.array(sha(np.bincount(r16.ravel().astype(np.int64), minlength=65536).astype(np.uint32)))
            x = frame.astype(np.float32) / 65535.0
            x3 = cv2.resize(np.stack([x] * 3, -1), (w, h))
            ti = torch.from_numpy(x3.transpose(2, 0, 1)).float()
            ei = ref.preprocessing.enhance_thermal_contrast(ti).numpy()
            out[tag + "_infer"] = np.array([sha(x3[..., 0]), sha(ei)])
            out[tag + "_infer_p"] = np.percentile(ti[0].numpy(), (2, 98))
            fr = ref.preprocessing.enhance_thermal_fixed_range(ti).numpy()
            out[tag + "_fixed"] = np.array([sha(fr), sha(ref.preprocessing.enhance_thermal_fixed_range(t, normalized=False).numpy())])
    # a small frame stored in full (inputs + outputs) so the GPU box can diff arrays, not only hashes
    small = ref_preprocess.make_raw_frames(1, seed=3, hw=(64, 80))[0]
    r16 = cv2.resize(small, (36, 28))
    t = torch.from_numpy(np.stack([r16.astype(np.float32)] * 3, 0)).float()
    out["small_raw"] = small
    out["small_resized_u16"] = r16
    out["small_train_out"] = ref.preprocessing.enhance_thermal_contrast(t).numpy()
    xs = cv2.resize(small.astype(np.float32) / 65535.0, (36, 28))
    out["small_resized_f32"] = xs
    out["small_infer_out"] = ref.preprocessing.enhance_thermal_contrast(torch.from_numpy(np.stack([xs] * 3, 0))).numpy()
    # non-collapsing 3-channel input (gray path) and a nearest resample
    rgbish = np.random.default_rng(5).random((3, 20, 24)).astype(np.float32)
    out["gray_in"] = rgbish
    out["gray_out"] = ref.preprocessing.enhance_thermal_contrast(torch.from_numpy(rgbish)).numpy()
    out["gray_fixed_out"] = ref.preprocessing.enhance_thermal_fixed_range(torch.from_numpy(rgbish)).numpy()
    d = np.random.default_rng(6).random((50, 70)).astype(np.float32)
    out["nearest_in"] = d
    out["nearest_out"] = cv2.resize(d, (33, 21), interpolation=cv2.INTER_NEAREST)
    np.savez_compressed(os.path.join(OUT, "preprocess_kat.npz"), **out)


def gen_metrics(ref):
    out = {"versions": versions(ref)}
    for split in ("day", "night"):
        files = sorted(glob.glob(os.path.join(ref.root, "pseudo_gt_test_set", split, "depth", "*_depth.npy")))
        gt_full, pred_full = np.load(files[0]), np.load(files[1])
        for ms in (True, False):
            m = ref.metrics.compute_depth_metrics(pred_full.copy(), gt_full.copy(), median_scaling=ms)
            out[f"{split}_full_ms{int(ms)}"] = np.array([m[k] for k in ("abs_rel", "sq_rel", "rmse", "rmse_log", "acc_1", "acc_2", "acc_3")], np.float64)
        # crops that travel to the GPU box
        gt, pred = gt_full[100:260, 200:392].copy(), pred_full[100:260, 200:392].copy()
        gt[10:20, 30:60] = 0.0            # invalid GT region
        gt[50, 5] = np.inf
        out[f"{split}_gt"], out[f"{split}_pred"] = gt, pred
        for ms in (True, False):
            m = ref.metrics.compute_depth_metrics(pred.copy(), gt.copy(), median_scaling=ms)
            out[f"{split}_crop_ms{int(ms)}"] = np.array([m[k] for k in ("abs_rel", "sq_rel", "rmse", "rmse_log", "acc_1", "acc_2", "acc_3")], np.float64)
            e = ref.evalm.compute_depth_metrics(pred.copy(), gt.copy(), median_scaling=ms)
            out[f"{split}_crop_eval_ms{int(ms)}"] = np.array([e["rmse"], e["acc_1.25"], e["acc_1.25^2"]], np.float64)
        mask = (np.random.default_rng(1).random(gt.shape) > 0.5) & (gt > 0) & np.isfinite(gt)
        m = ref.metrics.compute_depth_metrics(pred.copy(), gt.copy(), mask=mask)
        out[f"{split}_mask"] = mask
        out[f"{split}_crop_masked"] = np.array([m[k] for k in ("abs_rel", "sq_rel", "rmse", "rmse_log", "acc_1", "acc_2", "acc_3")], np.float64)
        # GT at another resolution, nearest-resampled as utils/evaluate_depth_metrics.py:320-323
        gt_big = gt_full[:300, :400].copy()
        gt_rs = ref.cv2.resize(gt_big, (pred.shape[1], pred.shape[0]), interpolation=ref.cv2.INTER_NEAREST)
        m = ref.metrics.compute_depth_metrics(pred.copy(), gt_rs, median_scaling=True)
        out[f"{split}_gt_big"] = gt_big
        out[f"{split}_crop_resampled"] = np.array([m[k] for k in ("abs_rel", "sq_rel", "rmse", "rmse_log", "acc_1", "acc_2", "acc_3")], np.float64)
    # calibration KATs + intrinsics estimate
    K, R, t = ref.pseudo_gt.load_thermal_calibration(os.path.join(ref.root, "calibrations", "t_calib.json"))
    Kl, Kr, T = ref.pseudo_gt.load_thermal_calibration(os.path.join(ref.root, "calibrations", "thermal_stereo_calib.yaml"))
    out["calib_json_K"], out["calib_json_R"], out["calib_json_t"] = K, R, t
    out["calib_yaml_Kl"], out["calib_yaml_Kr"], out["calib_yaml_T"] = Kl, Kr, T
    pm = np.random.default_rng(2).normal(size=(64, 80, 3)).astype(np.float32)
    pm[..., 2] = np.abs(pm[..., 2]) + 0.5
    pm[3, 4, 2] = -1.0
    out["focal_pointmap"] = pm
    out["focal_K"] = ref.pseudo_gt.estimate_camera_intrinsics(pm, pm[..., 2])
    np.savez_compressed(os.path.join(OUT, "metrics_kat.npz"), **out)


def gen_sobel(ref):
    m = ref.model.ThermalDUSt3R(torch.nn.Identity())
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    y = m.preprocess_thermal(x)
    xs = torch.rand(2, 1, 17, 23, generator=torch.Generator().manual_seed(1))
    m.edge_weight.data.fill_(0.8); m.temp_scale.data.fill_(0.9)
    xs_req = xs.clone().requires_grad_()
    ys = m.preprocess_thermal(xs_req)
    w = torch.rand(ys.shape, generator=torch.Generator().manual_seed(2))
    (ys * w).sum().backward()
    np.savez_compressed(os.path.join(OUT, "sobel_kat.npz"), big_sum=np.array(y.double().sum().item()),
                        big_sha=np.array(sha(y.detach().numpy())), small_in=xs.numpy(), small_out=ys.detach().numpy(),
                        small_w=w.numpy(), small_dx=xs_req.grad.numpy(), small_dew=m.edge_weight.grad.numpy(),
                        small_dts=m.temp_scale.grad.numpy(), versions=versions(ref))


def gen_evaluate(ref):
    """evaluate_thermal_depth (utils/metrics.py:72-138) on the fake model / loader of oracle/fake_eval.py."""
    from oracle import fake_eval
    out = {"versions": versions(ref)}
    for conv in fake_eval.CONVENTIONS:
        model = fake_eval.FakeModel(conv)
        res = ref.metrics.evaluate_thermal_depth(model, fake_eval.make_loader(seed=7), torch.device("cpu"))
        out[conv] = fake_eval.as_vector(res)
        out[conv + "_calls"] = np.array(model.calls)
        assert not model.training
    np.savez_compressed(os.path.join(OUT, "evaluate_kat.npz"), **out)


def gen_fire(ref):
    """Experimental fire-scene pipeline (thermal_dustr_inference_for_experiment.py:62-377): the LIVE reference functions
    (np.random seeded before each call) and the stock cv2 / numpy operators on the frames of ref_fire.make_fire_frame."""
    import importlib
    import sys
    import cv2
    from oracle import ref_fire
    sys.path.insert(0, ref.root)
    try:
        m = importlib.import_module("thermal_dustr_inference_for_experiment")
    finally:
        sys.path.remove(ref.root)
    out = {"versions": versions(ref), "sizes": np.array([[96, 128], [75, 100]])}
    for (h, w) in out["sizes"]:
        tag = f"{h}x{w}"
        f = ref_fire.make_fire_frame(int(h), int(w), seed=int(h))
        np.random.seed(11)
        out["pre_" + tag] = m.preprocess_fire_scene_thermal(torch.from_numpy(f)).numpy()
        np.random.seed(12)
        out["adv_" + tag] = m.advanced_fire_scene_processing(torch.from_numpy(f)).numpy()
        rng = np.random.default_rng(int(w))
        d = (3 + rng.standard_normal((int(h), int(w)))).astype(np.float32)
        d[rng.random(d.shape) < 0.01] += 30
        d[2:5, 3:6] += 40                                            # a clump: windows with few / no inliers
        out["depth_" + tag] = d
        out["refined_" + tag] = m.depth_refinement_with_outlier_removal(d.copy(), f, guided_filter=False)
        g = f[0]
        u8 = (g * 255).astype(np.uint8)
        out["u8_" + tag] = u8
        for clip in (2.5, 3.0):
            out[f"clahe{clip}_" + tag] = cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(u8)
        for lo in (30, 50):
            out[f"canny{lo}_" + tag] = cv2.Canny(u8, lo, 150)
        out["sobelx_" + tag] = cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=3)
        out["sobely_" + tag] = cv2.Sobel(g, cv2.CV_32F, 0, 1, ksize=3)
        out["hist_" + tag] = np.histogram(g.flatten(), bins=100, range=(0, 1))[0]
        out["bil5_" + tag] = cv2.bilateralFilter(d, 5, 50, 50)
        out["pct_" + tag] = np.array(np.percentile(g, (5, 95)))
    np.savez_compressed(os.path.join(OUT, "fire_kat.npz"), **out)


def main():
    ref = reference_bridge.load()
    os.makedirs(OUT, exist_ok=True)
    gen_loss(ref)
    gen_preprocess(ref)
    gen_metrics(ref)
    gen_sobel(ref)
    gen_evaluate(ref)
    gen_fire(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
