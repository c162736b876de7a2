"""utils.preprocessing drop-in (replaces /root/reference/utils/preprocessing.py)."""
from thermal3d_vision_b200.preprocessing import (enhance_thermal_contrast,  # noqa: F401
                                                 enhance_thermal_fixed_range)
