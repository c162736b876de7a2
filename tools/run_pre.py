"""Developer tool: run the preprocessing entry point a few times (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200 import preprocessing as pp, metrics as tm
from thermal3d_vision_b200.pipeline import HotPathStep
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev)
raw2 = torch.cat([d["raw1"], d["raw2"]])
hist = len(sys.argv) > 1 and sys.argv[1] == "hist"
with_loss = len(sys.argv) > 1 and sys.argv[1] == "loss"
args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
for _ in range(3):
    if with_loss:
        step.run_device(*args)
    else:
        pp.preprocess_thermal_batch(raw2, (W, H), out=step.pre_both, histogram=hist)
        tm.compute_depth_metrics_batch(d["pred1"], d["gt_depth"], out=step.met_out)
torch.cuda.synchronize()
