"""utils.metrics drop-in (replaces /root/reference/utils/metrics.py)."""
from thermal3d_vision_b200.metrics import compute_depth_metrics, evaluate_thermal_depth  # noqa: F401
