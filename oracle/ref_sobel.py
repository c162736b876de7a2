"""Oracle (test infrastructure, not product): Sobel thermal enhancer.

Restates /root/reference/thermal_dustr_model.py:110-142
(ThermalDUSt3R.preprocess_thermal) as a torch-CPU fp32 graph so autograd gives
d/d(edge_weight), d/d(temp_scale) and d/dx.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_SX = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]])
_SY = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]])


def preprocess_thermal_torch(x, edge_weight, temp_scale, local_norm=True):
    if x.size(1) == 1:                                        # :116-117
        x = x.repeat(1, 3, 1, 1)
    if local_norm:                                            # :120-124
        mn = x.amin(dim=(2, 3), keepdim=True)
        mx = x.amax(dim=(2, 3), keepdim=True)
        x = (x - mn) / (mx - mn + 1e-6)
    kx = _SX.to(x).reshape(1, 1, 3, 3).repeat(3, 1, 1, 1)
    ky = _SY.to(x).reshape(1, 1, 3, 3).repeat(3, 1, 1, 1)
    ex = F.conv2d(x, kx, padding=1, groups=3).abs()           # :131-132
    ey = F.conv2d(x, ky, padding=1, groups=3).abs()
    mag = torch.sqrt(ex.pow(2) + ey.pow(2))                   # :133
    out = (x + edge_weight * mag) * temp_scale                # :136-139
    return out.clamp(0, 1)                                    # :140
