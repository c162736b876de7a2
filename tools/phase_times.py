"""Developer tool: time the phases of HotPathStep.run_device with CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
from thermal3d_vision_b200 import preprocessing as pp, metrics as tm, loss as tl, _lib
dev = torch.device("cuda:0")
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, 0, dev)
step = HotPathStep(B, H, W, device=dev)
args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
res = {}
for pl in (True, False):
    step.pipelined = pl
    res[f"step_pipelined={pl}"] = timeit(lambda: step.run_device(*args))
    step.finish()
size = (W, H)
pre_half = [{}, {}]
res["pre1"] = timeit(lambda: pp.preprocess_thermal_batch(d["raw1"], size, out=pre_half[0]))
raw2 = torch.cat([d["raw1"], d["raw2"]])
out2 = {}
t = pp.preprocess_thermal_batch(raw2, size, out=out2)
out2.update({"thermal": t.thermal, "percentiles": t.percentiles, "histogram": t.histogram, "grad_stats": t.grad_stats})
res["pre_both_128"] = timeit(lambda: pp.preprocess_thermal_batch(raw2, size, out=out2))
res["metrics"] = timeit(lambda: tm.compute_depth_metrics_batch(d["pred1"], d["gt_depth"], out=step.met_out))
tb1 = pp.preprocess_thermal_batch(d["raw1"], size, out=pre_half[0]); tb2 = pp.preprocess_thermal_batch(d["raw2"], size, out=pre_half[1])
kw = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, multi_scale=False)
res["loss_with_stats"] = timeit(lambda: tl.fused_thermal_loss_fwd_bwd(d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], tb1.thermal, tb2.thermal, out=step.loss_out, thermal_stats=(tb1.grad_stats, tb2.grad_stats), **kw))
res["loss_with_stats_replicated"] = timeit(lambda: tl.fused_thermal_loss_fwd_bwd(d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], tb1.thermal, tb2.thermal, out=step.loss_out, thermal_stats=(tb1.grad_stats, tb2.grad_stats), thermal_replicated=True, **kw))
res["loss_no_stats"] = timeit(lambda: tl.fused_thermal_loss_fwd_bwd(d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], tb1.thermal, tb2.thermal, out=step.loss_out, **kw))
kwm = dict(kw); kwm["multi_scale"] = True
res["loss_multiscale"] = timeit(lambda: tl.fused_thermal_loss_fwd_bwd(d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], tb1.thermal, tb2.thermal, out=step.loss_out, **kwm))
# host-side enqueue cost of one step (no sync inside; the launch queue is deep enough for 20 steps)
import time
step.pipelined = True
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step.run_device(*args)
res["host_enqueue_us_per_step"] = (time.perf_counter() - t0) / 20 * 1e6
torch.cuda.synchronize()
res["loss_multiscale_replicated"] = timeit(lambda: tl.fused_thermal_loss_fwd_bwd(d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], tb1.thermal, tb2.thermal, out=step.loss_out, thermal_replicated=True, **kwm))
print(json.dumps(res, indent=1))
