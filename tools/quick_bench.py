"""Developer timing of individual entry points (CUDA events). Not the judged bench."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thermal3d_vision_b200 import loss as t3d, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64); ap.add_argument("--H", type=int, default=384)
ap.add_argument("--W", type=int, default=512); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--multi", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, H, W = a.B, a.H, a.W
g = torch.Generator(device=dev).manual_seed(0)
G1 = torch.randn(B, H, W, 3, device=dev, generator=g); G2 = torch.randn(B, H, W, 3, device=dev, generator=g)
P1 = G1 + 0.1 * torch.randn(B, H, W, 3, device=dev, generator=g); P2 = G2 + 0.1 * torch.randn(B, H, W, 3, device=dev, generator=g)
C1 = 1 + 4 * torch.rand(B, H, W, device=dev, generator=g); C2 = 1 + 4 * torch.rand(B, H, W, device=dev, generator=g)
T1 = torch.rand(B, 1, H, W, device=dev, generator=g).repeat(1, 3, 1, 1).contiguous()
T2 = torch.rand(B, 1, H, W, device=dev, generator=g).repeat(1, 3, 1, 1).contiguous()
out = {}
kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=bool(a.multi))
r = t3d.fused_thermal_loss_fwd_bwd(P1, P2, G1, G2, C1, C2, T1, T2, out=out, **kw)
out.update(r); out["workspace"] = torch.empty(_lib.lib().t3d_loss_workspace_bytes(B, H, W, 1), dtype=torch.uint8, device=dev)
for _ in range(3):
    t3d.fused_thermal_loss_fwd_bwd(P1, P2, G1, G2, C1, C2, T1, T2, out=out, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    t3d.fused_thermal_loss_fwd_bwd(P1, P2, G1, G2, C1, C2, T1, T2, out=out, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
bytes_alg = B * H * W * 112
print(json.dumps({"B": B, "H": H, "W": W, "multi": a.multi, "ms": ms, "pairs_per_s": B / ms * 1e3,
                  "alg_GBps": bytes_alg / ms / 1e6, "frac_of_6531.9": bytes_alg / ms / 1e6 / 6531.9,
                  "loss": out["batch"][0].item()}))
