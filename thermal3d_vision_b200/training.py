"""The loss-side glue of the reference's training / validation loops, batched and sync-free.

Mirrors train_thermal_dustr.py:214-293 (per-sample: resample the pseudo-GT to the prediction size
when they differ, choose the confidence -- predicted > ground-truth > ones --, clamp it to >= 1e-5,
evaluate the thermal-aware loss, keep finite positive samples, average) and :465-492 (validation:
plain mean-L1 over both views).  SURVEY.md section 8f rows 1-2.  The ViT forward is the caller's.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from . import loss as _loss


@_lib.on_tensor_device
def resample_bilinear(x: torch.Tensor, size_hw) -> torch.Tensor:
    """F.interpolate(mode='bilinear', align_corners=False) for channels-last maps.

    x: [B,H,W,3] pointmaps or [B,H,W] confidences (float32, CUDA) -> [B,h,w,(3)]."""
    squeeze = x.dim() == 3
    if squeeze:
        x = x.unsqueeze(-1)
    if x.dim() != 4:
        raise ValueError(f"expected [B,H,W] or [B,H,W,C], got {tuple(x.shape)}")
    _lib.require_cuda(x)
    x = x.float().contiguous()
    B, sh, sw, C = x.shape
    dh, dw = int(size_hw[0]), int(size_hw[1])
    out = torch.empty(B, dh, dw, C, dtype=torch.float32, device=x.device)
    rc = _lib.lib().t3d_interp_bilinear_f32(_lib.ptr(x), _lib.ptr(out), B, C, sh, sw, dh, dw, _lib.current_stream_ptr())
    _lib.check(rc, "t3d_interp_bilinear_f32")
    return out[..., 0] if squeeze else out


def training_batch_loss(pred_pts1, pred_pts2, gt_pts1, gt_pts2, pred_conf1=None, pred_conf2=None,
                        gt_conf1=None, gt_conf2=None, thermal1=None, thermal2=None, *, use_thermal_aware_loss=True,
                        alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=False,
                        thermal_stats=None):
    """One batched call for train_thermal_dustr.py:182-360's per-sample loop.

    Returns FusedLossResult (loss = mean over valid samples; gradients flow to pred_pts and, when they are
    the chosen confidence, to pred_conf).  Defaults are the CLI defaults (:52-55)."""
    # pseudo-GT at another resolution than the prediction (:234-271, the normal case: 512x512 vs 224x224): the
    # bilinear taps of the pointmaps -- and of the GT confidence when that is the one used -- are evaluated inside the
    # loss kernel's loads, the resampled arrays are never materialised (`resample_bilinear` is the stand-alone form)
    conf1 = pred_conf1 if pred_conf1 is not None else gt_conf1   # :274-275 (ones when both are None)
    conf2 = pred_conf2 if pred_conf2 is not None else gt_conf2
    if (conf1 is None) != (conf2 is None) or (conf1 is not None and conf1.shape != conf2.shape):
        H, W = pred_pts1.shape[1], pred_pts1.shape[2]            # mixed sources: bring the GT-sized one to the prediction's size
        conf1 = resample_bilinear(conf1, (H, W)) if conf1 is not None and tuple(conf1.shape[1:3]) != (H, W) else conf1
        conf2 = resample_bilinear(conf2, (H, W)) if conf2 is not None and tuple(conf2.shape[1:3]) != (H, W) else conf2
    # torch.clamp(conf, min=1e-5) (:278-279) composes idempotently with the loss's own clamp to [1e-5, 10]
    # (same values, same inclusive gradient mask), so it needs no extra pass.
    if not use_thermal_aware_loss:
        # :305-318: plain confidence-weighted L1 on the conf clamped from BELOW only (:278-279) -- it never passes
        # through utils/loss.py's clamp(conf, 1e-5, 10); DUSt3R confidences (1 + exp) routinely exceed 10
        return _loss.fused_thermal_loss(pred_pts1, pred_pts2, gt_pts1, gt_pts2, conf1, conf2, None, None,
                                        alpha=alpha, multi_scale=False, batch_mean=True, conf_min_only=True)
    return _loss.fused_thermal_loss(pred_pts1, pred_pts2, gt_pts1, gt_pts2, conf1, conf2, thermal1, thermal2,
                                    alpha=alpha, edge_weight=edge_weight, smoothness_weight=smoothness_weight,
                                    detail_weight=detail_weight, multi_scale=multi_scale, batch_mean=True)


def validation_batch_loss(pred_pts1, pred_pts2, gt_pts1, gt_pts2):
    """train_thermal_dustr.py:465-492: per sample (mean|p1-g1| + mean|p2-g2|) / 2, averaged over the finite,
    positive samples.  Returns FusedLossResult with that mean in `.loss`."""
    # (a GT of another size is resampled inside the kernel's loads, :465-481)
    # with confidence 1 and alpha 0 the basic term is mean|p1-g1| + mean|p2-g2| (utils/loss.py:81-98)
    r = _loss.fused_thermal_loss(pred_pts1, pred_pts2, gt_pts1, gt_pts2, None, None, None, None, alpha=0.0,
                                 multi_scale=False, batch_mean=True)
    return _loss.FusedLossResult(r.loss * 0.5, r.per_sample, r.batch)
