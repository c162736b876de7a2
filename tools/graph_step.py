"""Developer tool: HotPathStep as stream launches vs one CUDA-graph launch, at the launch-bound and the large shape."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
dev = torch.device("cuda:0")
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return round(best, 1)
res = {}
for (B, H, W) in ((8, 224, 224), (64, 384, 512)):
    d = bench.make_inputs_torch(B, H, W, 0, dev)
    step = HotPathStep(B, H, W, device=dev)
    args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
    res[f"B{B}_{W}x{H}_stream_us"] = timeit(lambda: step.run_device(*args))
    replay = step.capture_graph(*args)
    res[f"B{B}_{W}x{H}_graph_us"] = timeit(replay)
print(json.dumps(res))
