"""utils.loss drop-in (replaces /root/reference/utils/loss.py): same names and signatures."""
from thermal3d_vision_b200.loss import (confidence_weighted_regression_loss,  # noqa: F401
                                        enhanced_thermal_aware_loss, thermal_aware_loss)
