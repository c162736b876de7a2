"""Developer tool: kernel timeline of one EvalStep batch (BASELINE configs[4]) and the step time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import EvalStep
from thermal3d_vision_b200 import _lib
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H, W = 384, 512
step = EvalStep(B, H, W, gt_hw=(512, 512), device=dev)
g = torch.Generator(device=dev).manual_seed(1)
pool = []
for k in range(2):
    d = bench.make_inputs_torch(B, H, W, seed=k, device=dev)
    gt = 1.5 + 3 * torch.randn(B, 512, 512, device=dev, generator=g).abs()
    pm = d["pred1"].clone()
    pm[..., 2] = torch.nn.functional.interpolate(gt[:, None], size=(H, W), mode="nearest")[:, 0] * 0.7
    pool.append((d["raw1"], pm, gt)); del d
ahead = os.environ.get("T3D_EVAL_PREFETCH", "1") != "0"
def two():
    if ahead:
        step.prefetch(*pool[1]); step.run_batch(*pool[0]); step.prefetch(*pool[0]); step.run_batch(*pool[1])
    else:
        step.run_batch(*pool[0]); step.run_batch(*pool[1])
if ahead: step.prefetch(*pool[0])
ms = bench.time_steps(two, 10, 3) / 2
print(f"# eval step, batch {B}: {ms * 1e3:.1f} us  ({B / ms / 1e3 * 1e3:.0f} frames/s)")
_lib.profile_begin("", 256)
two(); torch.cuda.synchronize()
for nm, a, b in sorted(_lib.profile_timeline(), key=lambda r: r[1]): print(f"{a*1e3:9.1f} {b*1e3:9.1f} {(b-a)*1e3:7.1f}  {nm}")
_lib.profile_end()
