"""Developer tool: start/stop times of every kernel of a few HotPathStep calls (all streams), from CUDA events.

  python tools/timeline.py [steps] [plain|pipelined]

The bracketing events cost ~2 us of stream time per launch, so absolute times are a little longer than an
un-instrumented run; what the timeline shows is the ORDER and OVERLAP of the kernels of consecutive steps."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
from thermal3d_vision_b200 import _lib
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "pipelined"
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev, pipelined=(mode == "pipelined"))
args = tuple(d[k] for k in bench.KEYS)
for _ in range(10): step.run_device(*args)
step.finish(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): step.run_device(*args)
step.finish(); e1.record(); torch.cuda.synchronize()
print(f"# mode={mode}  un-instrumented: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per step")
_lib.profile_begin("", 4096)
for _ in range(nsteps): step.run_device(*args)
step.finish(); torch.cuda.synchronize()
tl = _lib.profile_timeline()
_lib.profile_end()
print("# start_us   stop_us   dur_us  kernel")
for nm, a, b in sorted(tl, key=lambda r: r[1]):
    print(f"{a*1e3:9.1f} {b*1e3:9.1f} {(b-a)*1e3:7.1f}  {nm}")
