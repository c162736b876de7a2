"""`thermal_dustr_model` drop-in (replaces /root/reference/thermal_dustr_model.py on PYTHONPATH).

`from thermal_dustr_model import load_dustr_model, ThermalDUSt3R` is what the reference's entry points do
(thermal_dustr_inference.py:21, utils/evaluate_depth_metrics.py:18, train_thermal_dustr.py:19).

* `ThermalDUSt3R` -- the Sobel thermal enhancer wrapper (thermal_dustr_model.py:86-200) -- resolves to the B200
  implementation (thermal3d_vision_b200.sobel: same constructor, parameters, state_dict keys, `preprocess_thermal`,
  `forward`, `save_checkpoint`; the arithmetic runs in libt3d_sm100.so).
* `load_dustr_model` builds the DUSt3R ViT (external naver/dust3r code, out of scope) and is NOT reimplemented: it is
  the reference's own function, loaded from the reference's file -- the first `thermal_dustr_model.py` other than
  this one found at $T3D_REFERENCE_ROOT or on sys.path.
"""
import importlib.util
import os
import sys

from thermal3d_vision_b200.sobel import ThermalDUSt3R, sobel_enhance  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_module():
    cands = [os.environ.get("T3D_REFERENCE_ROOT", "")] + list(sys.path)
    for c in cands:
        if not c:
            continue
        path = os.path.join(os.path.abspath(c), "thermal_dustr_model.py")
        if os.path.isfile(path) and os.path.dirname(path) != _HERE:
            spec = importlib.util.spec_from_file_location("_t3d_reference_thermal_dustr_model", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


_ref = _reference_module()
if _ref is not None:
    load_dustr_model = _ref.load_dustr_model
else:
    def load_dustr_model(weights_path, device=None, is_thermal=False):
        raise ImportError("load_dustr_model is the reference's own function (it builds the external DUSt3R ViT): put the "
                          "Thermal3D-Vision checkout on PYTHONPATH after this drop-in, or set T3D_REFERENCE_ROOT")
