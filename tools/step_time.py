"""Developer tool: device time of one HotPathStep (overlap on), for env-var sweeps."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev, pipelined=bool(int(os.environ.get("T3D_PIPELINED", "1"))))
args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
main_prio = int(os.environ.get("T3D_MAIN_PRIO", "0"))
ms = torch.cuda.Stream(device=dev, priority=main_prio)
torch.cuda.set_stream(ms)
for _ in range(10): step.run_device(*args)
step.finish()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(100): step.run_device(*args)
    step.finish()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100 * 1e3)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("T3D_")}, "step_us": round(best, 1)}))
