"""Developer tool: per-kernel CUDA-event times (t3d_profile_begin/end) of the preprocessing and metric entry points."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200 import preprocessing as pp, metrics as tm, _lib
from thermal3d_vision_b200.pipeline import HotPathStep
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev)
raw2 = torch.cat([d["raw1"], d["raw2"]])
res = {}
def per_kernel(names, fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    out = {}
    for nm in names:
        _lib.profile_begin(nm, 4096)
        for _ in range(iters): fn()
        torch.cuda.synchronize()
        ms, n = _lib.profile_end()
        out[nm] = round(ms / iters * 1e3, 2)     # us per call (all launches of that kernel in one call)
    return out
for hist in (True, False):
    fn = lambda: pp.preprocess_thermal_batch(raw2, (W, H), out=step.pre_both, histogram=hist)
    names = ["build_taps", "resize_hist", "percentile_from_hist", "normalize_stats"] if hist else \
            ["build_taps", "bracket_sample", "resize_march", "percentile_from_brackets", "normalize_stats"]
    r = per_kernel(names, fn)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    r["total_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 2)
    res[f"preprocess_128_hist={hist}"] = r
fn = lambda: tm.compute_depth_metrics_batch(d["pred1"], d["gt_depth"], out=step.met_out)
r = per_kernel(["metrics_sample", "depth_extract", "median_scale", "metrics_sum", "metrics_finalize"], fn)
res["metrics_64"] = r
print(json.dumps(res, indent=1))
