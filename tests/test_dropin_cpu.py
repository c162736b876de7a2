"""CPU: the drop-in `utils` shim resolves the replaced modules to our implementation (same public
names and signatures as the reference) and everything else to the reference."""
import inspect
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SIGS = {   # positional order, keyword names, defaults
    "loss.confidence_weighted_regression_loss": "(pred_pts1, pred_pts2, gt_pts1, gt_pts2, confidences1=None, confidences2=None, alpha=0.2)",
    "loss.thermal_aware_loss": "(pred_pts1, pred_pts2, gt_pts1, gt_pts2, confidences1=None, confidences2=None, thermal_img1=None, thermal_img2=None, alpha=0.2, edge_weight=0.5, smoothness_weight=0.3)",
    "loss.enhanced_thermal_aware_loss": "(pred_pts1, pred_pts2, gt_pts1, gt_pts2, confidences1=None, confidences2=None, thermal_img1=None, thermal_img2=None, alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.3, multi_scale=True)",
    "preprocessing.enhance_thermal_contrast": "(thermal_tensor)",
    "preprocessing.enhance_thermal_fixed_range": "(thermal_tensor, normalized=True)",
    "metrics.compute_depth_metrics": "(pred_depth, gt_depth, mask=None, median_scaling=True)",
    "metrics.evaluate_thermal_depth": "(model, dataloader, device)",
    "preprocessing.load_and_preprocess_thermal_image": "(path, img_size=(224, 224))",
    "depth.load_thermal_calibration": "(calib_path)",
    "depth.estimate_camera_intrinsics": "(pointmap, depth, calib_path=None)",
}


def test_signatures_match_the_reference():
    import importlib
    for name, sig in SIGS.items():
        mod, fn = name.split(".")
        f = getattr(importlib.import_module(f"thermal3d_vision_b200.{mod}"), fn)
        assert str(inspect.signature(f)) == sig, name


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_signatures_equal_live_reference():
    from oracle import reference_bridge
    ref = reference_bridge.load()
    import thermal3d_vision_b200.loss as L, thermal3d_vision_b200.metrics as M, thermal3d_vision_b200.preprocessing as P
    pairs = [(L.enhanced_thermal_aware_loss, ref.loss.enhanced_thermal_aware_loss),
             (L.thermal_aware_loss, ref.loss.thermal_aware_loss),
             (L.confidence_weighted_regression_loss, ref.loss.confidence_weighted_regression_loss),
             (P.enhance_thermal_contrast, ref.preprocessing.enhance_thermal_contrast),
             (P.enhance_thermal_fixed_range, ref.preprocessing.enhance_thermal_fixed_range),
             (M.compute_depth_metrics, ref.metrics.compute_depth_metrics),
             (M.evaluate_thermal_depth, ref.metrics.evaluate_thermal_depth),
             (P.load_and_preprocess_thermal_image, ref.evalm.load_and_preprocess_thermal_image)]
    for ours, theirs in pairs:
        assert str(inspect.signature(ours)) == str(inspect.signature(theirs)), ours.__name__


def test_shim_resolution():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, REF]), PYTHONDONTWRITEBYTECODE="1")
    code = ("import utils.loss, utils.preprocessing, utils.metrics;"
            "print(utils.loss.enhanced_thermal_aware_loss.__module__);"
            "print(utils.preprocessing.enhance_thermal_contrast.__module__);"
            "print(utils.metrics.compute_depth_metrics.__module__)")
    code += (";import thermal_dustr_model as m; print(m.ThermalDUSt3R.__module__); print(m.load_dustr_model.__module__);"
             "import inspect; print(inspect.signature(m.load_dustr_model))")
    if os.path.isdir(REF):
        code += ";import utils.data_utils; print(utils.data_utils.__file__)"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "thermal3d_vision_b200.loss" and lines[1] == "thermal3d_vision_b200.preprocessing"
    assert lines[2] == "thermal3d_vision_b200.metrics"
    # thermal_dustr_inference.py:21,95-96: ThermalDUSt3R is ours, load_dustr_model stays the reference's own function
    assert lines[3] == "thermal3d_vision_b200.sobel"
    assert lines[5] == "(weights_path, device=None, is_thermal=False)"
    if os.path.isdir(REF):
        assert lines[4] == "_t3d_reference_thermal_dustr_model"
        assert lines[6].startswith(REF)
