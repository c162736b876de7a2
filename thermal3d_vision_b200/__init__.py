"""thermal3d_vision_b200 -- B200-native (sm_100a) implementation of the
Thermal3D-Vision per-pixel hot path: thermal preprocessing, the thermal-aware
training loss (fused forward+backward), pointmap->depth and depth metrics.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all
arithmetic runs in hand-written CUDA behind the C ABI of ``libt3d_sm100.so``
(``include/t3d.h``).  There is no CPU fallback and no Triton.
"""
from . import _lib
from ._lib import T3DError, build, launch_count

__all__ = ["_lib", "T3DError", "build", "launch_count"]
__version__ = "0.1.0"
