"""Developer tool: start/stop times of every kernel of a few HotPathStep calls (all streams), from CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
from thermal3d_vision_b200 import _lib
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev)
args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
for _ in range(10): step.run_device(*args)
torch.cuda.synchronize()
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
_lib.profile_begin("", 4096)
for _ in range(nsteps): step.run_device(*args)
torch.cuda.synchronize()
tl = _lib.profile_timeline()
_lib.profile_end()
for nm, a, b in sorted(tl, key=lambda r: r[1]):
    print(f"{a*1e3:9.1f} {b*1e3:9.1f} {(b-a)*1e3:7.1f}  {nm}")
