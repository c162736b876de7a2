"""Developer tool: time of preprocess_thermal_batch(128 frames, histogram=False) for env-var sweeps."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200 import preprocessing as pp
from thermal3d_vision_b200.pipeline import HotPathStep
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev)
raw2 = torch.cat([d["raw1"], d["raw2"]])
fn = lambda: pp.preprocess_thermal_batch(raw2, (W, H), out=step.pre_both, histogram=False)
for _ in range(5): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(3):
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 50 * 1e3)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("T3D_")}, "pre_us": round(best, 1)}))
