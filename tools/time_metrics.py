"""Developer timing of the metrics / preprocess entry points."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thermal3d_vision_b200 import metrics as tm, preprocessing as pp
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
g = torch.Generator(device=dev).manual_seed(0)
gt = 1.5 + 3 * torch.randn(B, H, W, device=dev, generator=g).abs()
pm = torch.randn(B, H, W, 3, device=dev, generator=g); pm[..., 2] = gt * (1 + 0.1 * torch.randn(B, H, W, device=dev, generator=g))
raw = (22800 + 400 * torch.randn(B, 512, 640, device=dev, generator=g)).clamp(0, 65535).to(torch.int32).to(torch.uint16)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out_m, out_p = {}, {}
r = tm.compute_depth_metrics_batch(pm, gt, out=out_m); out_m.update(r)
t = pp.preprocess_thermal_batch(raw, (W, H), out=out_p)
out_p.update({"thermal": t.thermal, "percentiles": t.percentiles, "histogram": t.histogram, "grad_stats": t.grad_stats})
print(json.dumps({"metrics_us": timeit(lambda: tm.compute_depth_metrics_batch(pm, gt, out=out_m)),
                  "preprocess_us": timeit(lambda: pp.preprocess_thermal_batch(raw, (W, H), out=out_p)),
                  "chunk_px": os.environ.get("T3D_METRIC_CHUNK_PX")}))
