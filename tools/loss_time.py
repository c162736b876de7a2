"""Developer tool: device time of the loss call alone (march kernel + second stage), for env-var sweeps."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
from thermal3d_vision_b200 import preprocessing as pp, loss as tl, _lib
dev = torch.device("cuda:0")
MS = len(sys.argv) > 1 and sys.argv[1] == "ms"       # multi-scale loss (reference default)
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev)
raw2 = torch.cat([d["raw1"], d["raw2"]])
tb = pp.preprocess_thermal_batch(raw2, (W, H), out=step.pre_both, histogram=False)
kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=MS)
fn = lambda: tl.fused_thermal_loss_fwd_bwd(d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], tb.thermal[:B], tb.thermal[B:],
                                           out=step.loss_out, thermal_stats=None if MS else (tb.grad_stats[:B], tb.grad_stats[B:]), thermal_replicated=True,
                                           rescale_invalid=False, **kw)
for _ in range(5): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50 * 1e3)
_lib.profile_begin("loss_march", 64)       # matches loss_march_kernel and loss_march_ms_kernel
for _ in range(20): fn()
torch.cuda.synchronize()
ms, n = _lib.profile_end()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("T3D_")}, "loss_call_us": round(best, 1), "march_kernel_us": round(ms / n * 1e3, 1)}))
