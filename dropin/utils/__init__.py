"""Drop-in `utils` package for the reference's entry points.

Put this directory BEFORE the reference checkout on PYTHONPATH:

    PYTHONPATH=/path/to/thermal3d-b200/dropin:/path/to/thermal3d-b200:/path/to/Thermal3D-Vision \
        python /path/to/Thermal3D-Vision/train_thermal_dustr.py ...

`utils.loss`, `utils.preprocessing`, `utils.metrics` resolve to the B200 implementation
(thermal3d_vision_b200); every other `utils.*` module (visualize, data_utils, ...) still resolves
to the reference's own file, found next to a `train_thermal_dustr.py` on sys.path or at
$T3D_REFERENCE_ROOT.
"""
import os
import sys


def _reference_utils_dir():
    cands = [os.environ.get("T3D_REFERENCE_ROOT", "")] + list(sys.path)
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "train_thermal_dustr.py")) and \
                os.path.isdir(os.path.join(c, "utils")):
            return os.path.join(c, "utils")
    return None


_ref = _reference_utils_dir()
if _ref and _ref not in __path__:
    __path__.append(_ref)        # fall through to the reference for modules we do not replace
