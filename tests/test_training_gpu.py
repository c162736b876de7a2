"""GPU parity for the SURVEY.md 8f 'next' rows: GT/conf bilinear resampling (F.interpolate semantics) and the
batched training / validation loss wrappers, against the reference's per-sample loop restated with torch."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_loss

pytestmark = pytest.mark.gpu
KW = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=False)


def _interp_like_reference(x, size):
    """train_thermal_dustr.py:239-268: [H,W,3] -> permute -> F.interpolate -> permute back; [H,W] conf alike."""
    if x.dim() == 4:
        return F.interpolate(x.permute(0, 3, 1, 2), size=size, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    return F.interpolate(x.unsqueeze(1), size=size, mode="bilinear", align_corners=False).squeeze(1)


@pytest.mark.parametrize("src,dst", [((64, 64), (28, 28)), ((48, 80), (56, 56)), ((33, 47), (33, 47)), ((17, 23), (40, 31))])
def test_resample_matches_f_interpolate(cuda_device, src, dst):
    from thermal3d_vision_b200.training import resample_bilinear
    g = torch.Generator().manual_seed(src[0] * 100 + dst[1])
    pm = torch.randn(2, *src, 3, generator=g)
    cf = torch.rand(2, *src, generator=g)
    torch.testing.assert_close(resample_bilinear(pm.to(cuda_device), dst).cpu(), _interp_like_reference(pm, dst),
                               rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(resample_bilinear(cf.to(cuda_device), dst).cpu(), _interp_like_reference(cf, dst),
                               rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("conf_source", ["pred", "gt", "none"])
def test_training_batch_loss_matches_reference_loop(cuda_device, conf_source):
    from thermal3d_vision_b200.training import training_batch_loss
    B, H, W, GH, GW = 3, 40, 48, 64, 64
    P1, P2, _, _, C1, C2, T1, T2 = ref_loss.make_batch_inputs(B, H, W, seed=2)
    g = torch.Generator().manual_seed(5)
    G1 = torch.randn(B, GH, GW, 3, generator=g); G1[..., 2] = 1.5 + 3 * G1[..., 2].abs()
    G2 = torch.randn(B, GH, GW, 3, generator=g); G2[..., 2] = 1.5 + 3 * G2[..., 2].abs()
    GC1, GC2 = 1 + 4 * torch.rand(B, GH, GW, generator=g), 1 + 4 * torch.rand(B, GH, GW, generator=g)

    # reference loop (train_thermal_dustr.py:182-360) on the CPU
    p1, p2 = P1.clone().requires_grad_(), P2.clone().requires_grad_()
    c1, c2 = C1.clone().requires_grad_(), C2.clone().requires_grad_()
    tot, nv = 0.0, 0
    for i in range(B):
        g1, g2 = _interp_like_reference(G1[i:i + 1], (H, W))[0], _interp_like_reference(G2[i:i + 1], (H, W))[0]
        gc1, gc2 = _interp_like_reference(GC1[i:i + 1], (H, W))[0], _interp_like_reference(GC2[i:i + 1], (H, W))[0]
        cf1 = c1[i] if conf_source == "pred" else (gc1 if conf_source == "gt" else torch.ones(H, W))
        cf2 = c2[i] if conf_source == "pred" else (gc2 if conf_source == "gt" else torch.ones(H, W))
        loss, _ = ref_loss.enhanced_thermal_aware_loss_torch(p1[i], p2[i], g1, g2, torch.clamp(cf1, min=1e-5),
                                                             torch.clamp(cf2, min=1e-5), T1[i], T2[i], **KW)
        if torch.isfinite(loss) and loss > 0:
            tot = tot + loss; nv += 1
    tot = tot / nv
    tot.backward()

    d = lambda t: t.to(cuda_device)
    q1, q2 = d(P1).requires_grad_(), d(P2).requires_grad_()
    k1, k2 = d(C1).requires_grad_(), d(C2).requires_grad_()
    res = training_batch_loss(q1, q2, d(G1), d(G2),
                              pred_conf1=k1 if conf_source == "pred" else None, pred_conf2=k2 if conf_source == "pred" else None,
                              gt_conf1=d(GC1) if conf_source == "gt" else None, gt_conf2=d(GC2) if conf_source == "gt" else None,
                              thermal1=d(T1), thermal2=d(T2))
    res.loss.backward()
    assert res.loss.item() == pytest.approx(tot.item(), rel=1e-5)
    torch.testing.assert_close(q1.grad.cpu(), p1.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(q2.grad.cpu(), p2.grad, rtol=1e-4, atol=1e-6)
    if conf_source == "pred":
        torch.testing.assert_close(k1.grad.cpu(), c1.grad, rtol=1e-4, atol=1e-6)


def test_plain_loss_path_clamps_confidence_from_below_only(cuda_device):
    """use_thermal_aware_loss=False (train_thermal_dustr.py:305-318): confidences above 10 (DUSt3R's 1 + exp routinely
    is) are NOT clamped, unlike utils/loss.py:91-92; below 1e-5 they are (and take no gradient)."""
    from thermal3d_vision_b200.training import training_batch_loss
    B, H, W = 3, 36, 44
    P1, P2, G1, G2, *_ = ref_loss.make_batch_inputs(B, H, W, seed=4)
    g = torch.Generator().manual_seed(9)
    C1 = 30 * torch.rand(B, H, W, generator=g) - 1.0        # spans < 1e-5 (incl. negative) ... 29
    C2 = 1 + torch.exp(3 * torch.randn(B, H, W, generator=g))
    p1, p2 = P1.clone().requires_grad_(), P2.clone().requires_grad_()
    c1, c2 = C1.clone().requires_grad_(), C2.clone().requires_grad_()
    tot, nv = 0.0, 0
    for i in range(B):
        loss = ref_loss.plain_confidence_loss_torch(p1[i], p2[i], G1[i], G2[i], c1[i], c2[i])
        if torch.isfinite(loss) and loss > 0:
            tot = tot + loss; nv += 1
    assert nv == B and (C1 > 10).any() and (C2 > 10).any() and (C1 < 1e-5).any()
    (tot / nv).backward()
    d = lambda t: t.to(cuda_device)
    q1, q2, k1, k2 = d(P1).requires_grad_(), d(P2).requires_grad_(), d(C1).requires_grad_(), d(C2).requires_grad_()
    res = training_batch_loss(q1, q2, d(G1), d(G2), pred_conf1=k1, pred_conf2=k2, use_thermal_aware_loss=False)
    res.loss.backward()
    assert res.loss.item() == pytest.approx((tot / nv).item(), rel=1e-5)
    for got, ref in ((q1.grad, p1.grad), (q2.grad, p2.grad), (k1.grad, c1.grad), (k2.grad, c2.grad)):
        torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-6)
    # the thermal-aware path keeps utils/loss.py's clamp to [1e-5, 10]: a different number on the same inputs
    clamped = training_batch_loss(d(P1), d(P2), d(G1), d(G2), pred_conf1=d(C1), pred_conf2=d(C2), thermal1=None, thermal2=None)
    assert abs(clamped.loss.item() - res.loss.item()) > 1e-3


@pytest.mark.parametrize("multi", [False, True])
@pytest.mark.parametrize("H,W,GH,GW", [(224, 224, 512, 512), (37, 50, 64, 48), (48, 64, 24, 40)])
def test_fused_gt_resampling_equals_resample_then_loss(cuda_device, H, W, GH, GW, multi):
    """The pseudo-GT at another resolution (the normal case: 512x512 vs 224x224, train_thermal_dustr.py:234-271): the
    taps fused into the loss kernel's loads give what resampling first (F.interpolate semantics, checked above) and
    then evaluating the loss gives -- loss, components and all gradients -- and both match the reference loop."""
    from thermal3d_vision_b200 import loss as t3d
    from thermal3d_vision_b200.training import resample_bilinear
    B = 2
    P1, P2, _, _, C1, C2, T1, T2 = ref_loss.make_batch_inputs(B, H, W, seed=H + GW)
    g = torch.Generator().manual_seed(GH)
    G1 = torch.randn(B, GH, GW, 3, generator=g); G1[..., 2] = 1.5 + 3 * G1[..., 2].abs()
    G2 = torch.randn(B, GH, GW, 3, generator=g); G2[..., 2] = 1.5 + 3 * G2[..., 2].abs()
    GC1, GC2 = 1 + 4 * torch.rand(B, GH, GW, generator=g), 1 + 4 * torch.rand(B, GH, GW, generator=g)
    d = lambda t: t.to(cuda_device)
    kw = dict(KW, multi_scale=multi)
    for conf_src in ("pred", "gt"):
        cA, cB = (C1, C2) if conf_src == "pred" else (GC1, GC2)
        q = [d(P1).requires_grad_(), d(P2).requires_grad_()]
        k = [d(cA).requires_grad_(conf_src == "pred"), d(cB).requires_grad_(conf_src == "pred")]
        fused = t3d.fused_thermal_loss(q[0], q[1], d(G1), d(G2), k[0], k[1], d(T1), d(T2), **kw)
        fused.loss.backward()
        q2 = [d(P1).requires_grad_(), d(P2).requires_grad_()]
        c2 = [d(cA), d(cB)] if conf_src == "pred" else [resample_bilinear(d(cA), (H, W)), resample_bilinear(d(cB), (H, W))]
        k2 = [c.requires_grad_(conf_src == "pred") for c in c2]
        two = t3d.fused_thermal_loss(q2[0], q2[1], resample_bilinear(d(G1), (H, W)), resample_bilinear(d(G2), (H, W)),
                                     k2[0], k2[1], d(T1), d(T2), **kw)
        two.loss.backward()
        torch.testing.assert_close(fused.per_sample[:, :5], two.per_sample[:, :5], rtol=2e-6, atol=1e-7)
        torch.testing.assert_close(q[0].grad, q2[0].grad, rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(q[1].grad, q2[1].grad, rtol=1e-4, atol=1e-7)
        if conf_src == "pred":
            torch.testing.assert_close(k[0].grad, k2[0].grad, rtol=1e-4, atol=1e-7)
    # against the reference arithmetic on the CPU (resample with F.interpolate, loss per sample)
    p1 = P1.clone().requires_grad_()
    tot = 0.0
    for i in range(B):
        g1, g2 = _interp_like_reference(G1[i:i + 1], (H, W))[0], _interp_like_reference(G2[i:i + 1], (H, W))[0]
        loss, _ = ref_loss.enhanced_thermal_aware_loss_torch(p1[i], P2[i], g1, g2, C1[i], C2[i], T1[i], T2[i], **kw)
        tot = tot + loss
    (tot / B).backward()
    q = d(P1).requires_grad_()
    r = t3d.fused_thermal_loss(q, d(P2), d(G1), d(G2), d(C1), d(C2), d(T1), d(T2), **kw)
    r.loss.backward()
    assert r.loss.item() == pytest.approx((tot / B).item(), rel=1e-5)
    torch.testing.assert_close(q.grad.cpu(), p1.grad, rtol=1e-4, atol=1e-6)


def test_validation_batch_loss(cuda_device):
    from thermal3d_vision_b200.training import validation_batch_loss
    B, H, W = 4, 32, 36
    P1, P2, G1, G2, *_ = ref_loss.make_batch_inputs(B, H, W, seed=8)
    ref = sum(((P1[i] - G1[i]).abs().mean(-1).mean() + (P2[i] - G2[i]).abs().mean(-1).mean()) / 2 for i in range(B)) / B
    got = validation_batch_loss(*(t.to(cuda_device) for t in (P1, P2, G1, G2)))
    assert got.loss.item() == pytest.approx(ref.item(), rel=1e-5)
