"""Oracle helper (test infrastructure): import the UNMODIFIED reference.

Only works where /root/reference is mounted (the build container).  Used by
``oracle/gen_golden.py`` and ``tests/test_oracle_pin.py``; never on the GPU
box, never by the product.  Recipe: SURVEY.md Appendix E (matplotlib stubs,
IPP off so cv2.resize is the deterministic C++ path).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("T3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utils", "loss.py"))


_loaded = None


def load():
    """Return a namespace with the reference's hot-path callables."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits",
                 "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

    # The reference is a flat script repo whose packages are called `utils`
    # and `data`; import them under a private sys.path and restore afterwards
    # so a drop-in `utils` shim on the path is not shadowed.
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                  if k == "utils" or k.startswith("utils.")}
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "scripts"))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import cv2
        cv2.ipp.setUseIPP(False)
        import utils.loss as rloss
        import utils.preprocessing as rpre
        import utils.metrics as rmet
        import utils.evaluate_depth_metrics as reval
        import thermal_dustr_model as rmodel
        import pseudo_gt as rpgt
    finally:
        sys.path[:] = saved_path
        ref_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                    if k == "utils" or k.startswith("utils.")}
        sys.modules.update(saved_mods)
    ns = types.SimpleNamespace(
        loss=rloss, preprocessing=rpre, metrics=rmet, evalm=reval, model=rmodel, pseudo_gt=rpgt,
        cv2=cv2, root=REFERENCE_ROOT, _mods=ref_mods)
    _loaded = ns
    return ns
