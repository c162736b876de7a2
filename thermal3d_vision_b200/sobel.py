"""Sobel thermal enhancer -- host side of csrc/t3d_sobel.cu.

Mirrors /root/reference/thermal_dustr_model.py:86-200 (class ThermalDUSt3R): same
constructor, parameters (edge_weight 0.5, temp_scale 1.0), `preprocess_thermal`,
`forward` and `save_checkpoint`.  The DUSt3R ViT it wraps is the caller's (out of scope).
Gradients flow to the two learnable scalars; the thermal input itself takes no gradient
(in the reference it never requires one, and sqrt'(0) makes that gradient NaN on flat regions).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class _SobelEnhance(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, edge_weight, temp_scale, local_norm):
        B, C, H, W = x.shape
        lib = _lib.lib()
        params = torch.stack([edge_weight.detach().reshape(()), temp_scale.detach().reshape(())]).to(
            device=x.device, dtype=torch.float32).contiguous()
        ws = torch.empty(lib.t3d_sobel_workspace_bytes(B, C, H, W), dtype=torch.uint8, device=x.device)
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
        rc = lib.t3d_sobel_enhance_fwd(_lib.ptr(x), _lib.ptr(params), B, C, H, W, int(local_norm), _lib.ptr(out),
                                       _lib.ptr(ws), ws.numel(), _lib.current_stream_ptr())
        _lib.check(rc, "t3d_sobel_enhance_fwd")
        ctx.save_for_backward(x, params)
        ctx.local_norm = bool(local_norm)
        return out

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, dout):
        x, params = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise _lib.T3DError("gradient w.r.t. the thermal input is not provided (NaN-prone in the reference)")
        B, C, H, W = x.shape
        lib = _lib.lib()
        ws = torch.empty(lib.t3d_sobel_workspace_bytes(B, C, H, W), dtype=torch.uint8, device=x.device)
        dparams = torch.empty(2, dtype=torch.float32, device=x.device)
        rc = lib.t3d_sobel_enhance_bwd_params(_lib.ptr(x), _lib.ptr(params), _lib.ptr(dout.contiguous().float()),
                                              B, C, H, W, int(ctx.local_norm), _lib.ptr(dparams), _lib.ptr(ws),
                                              ws.numel(), _lib.current_stream_ptr())
        _lib.check(rc, "t3d_sobel_enhance_bwd_params")
        return None, dparams[0], dparams[1], None


def sobel_enhance(x: torch.Tensor, edge_weight, temp_scale, use_local_normalization: bool = True) -> torch.Tensor:
    """Functional form of ThermalDUSt3R.preprocess_thermal: x [B,1|3,H,W] -> [B,3,H,W]."""
    src_cuda = x.is_cuda
    if not src_cuda:
        if not torch.cuda.is_available():
            raise _lib.T3DError("no CUDA device: thermal3d_vision_b200 has no CPU path")
        x = x.cuda()
    x = x.float().contiguous()
    if x.dim() != 4 or x.shape[1] not in (1, 3):
        raise ValueError(f"expected [B,1|3,H,W], got {tuple(x.shape)}")
    ew = edge_weight if isinstance(edge_weight, torch.Tensor) else torch.tensor(float(edge_weight))
    ts = temp_scale if isinstance(temp_scale, torch.Tensor) else torch.tensor(float(temp_scale))
    out = _SobelEnhance.apply(x, ew, ts, use_local_normalization)
    return out if src_cuda else out.cpu()


class ThermalDUSt3R(nn.Module):
    """Drop-in for thermal_dustr_model.py:86-200."""

    def __init__(self, base_model):
        super().__init__()
        self.model = base_model
        sobel_x = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=torch.float32).reshape(1, 1, 3, 3)
        sobel_y = torch.tensor([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], dtype=torch.float32).reshape(1, 1, 3, 3)
        self.register_buffer("sobel_x", sobel_x.repeat(3, 1, 1, 1))      # kept for state_dict compatibility
        self.register_buffer("sobel_y", sobel_y.repeat(3, 1, 1, 1))
        self.edge_weight = nn.Parameter(torch.tensor(0.5))
        self.temp_scale = nn.Parameter(torch.tensor(1.0))
        self.use_local_normalization = True

    def preprocess_thermal(self, x):
        return sobel_enhance(x, self.edge_weight, self.temp_scale, self.use_local_normalization)

    def forward(self, view1, view2):
        if isinstance(view1, dict) and "img" in view1:
            v1, v2 = view1.copy(), view2.copy()
            v1["img"] = self.preprocess_thermal(view1["img"])
            v2["img"] = self.preprocess_thermal(view2["img"])
            return self.model(v1, v2)
        return self.model(self.preprocess_thermal(view1), self.preprocess_thermal(view2))

    def save_checkpoint(self, path, optimizer=None, epoch=None, val_loss=None, args=None):
        """Same checkpoint layout as thermal_dustr_model.py:191-200."""
        torch.save({"epoch": epoch, "state_dict": self.state_dict(),
                    "optimizer": None if optimizer is None else optimizer.state_dict(),
                    "val_loss": val_loss, "args": args}, path)
