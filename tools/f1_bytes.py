"""Developer tool: SURVEY 8f row 1 -- pseudo-GT at 512x512 against 224x224 predictions (batch 8): fused taps in the loss
kernel's loads vs the stand-alone resample kernel + loss.  Time per call; run under ncu for the DRAM bytes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200 import loss as tl
from thermal3d_vision_b200.training import resample_bilinear
dev = torch.device("cuda:0")
B, H, W, GH, GW = 8, 224, 224, 512, 512
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
G1, G2 = r(B, GH, GW, 3), r(B, GH, GW, 3)
G1[..., 2] = 1.5 + 3 * G1[..., 2].abs(); G2[..., 2] = 1.5 + 3 * G2[..., 2].abs()
P1, P2 = r(B, H, W, 3), r(B, H, W, 3)
C1 = 1 + 4 * torch.rand(B, H, W, device=dev, generator=g); C2 = 1 + 4 * torch.rand(B, H, W, device=dev, generator=g)
T = torch.rand(B, 1, H, W, device=dev, generator=g).repeat(1, 3, 1, 1).contiguous()
kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=False)
out_f, out_s = {}, {}
def fused():
    out_f.update(tl.fused_thermal_loss_fwd_bwd(P1, P2, G1, G2, C1, C2, T, T, out=out_f, **kw))
def separate():
    g1, g2 = resample_bilinear(G1, (H, W)), resample_bilinear(G2, (H, W))
    out_s.update(tl.fused_thermal_loss_fwd_bwd(P1, P2, g1, g2, C1, C2, T, T, out=out_s, **kw))
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
if mode == "time":
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    res = {}
    for name, fn in (("fused", fused), ("separate", separate)):
        for _ in range(3): fn()
        ts = []
        for _ in range(20):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        res[name + "_us"] = sorted(ts)[len(ts) // 2]
    res["loss_equal_rel"] = abs(out_f["batch"][0].item() - out_s["batch"][0].item()) / abs(out_s["batch"][0].item())
    print(json.dumps(res))
else:
    {"fused": fused, "separate": separate}[mode](); torch.cuda.synchronize()
    {"fused": fused, "separate": separate}[mode](); torch.cuda.synchronize()
