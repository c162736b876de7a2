"""Developer tool: inference-path preprocessing times."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200 import preprocessing as pp
dev = torch.device("cuda:0")
d = bench.make_inputs_torch(64, 384, 512, 0, dev)
raw = d["raw1"]
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 1)
res = {}
res["train_64_us"] = timeit(lambda: pp.preprocess_thermal_batch(raw, (512, 384), path="train", histogram=False))
res["inference_64_us"] = timeit(lambda: pp.preprocess_thermal_batch(raw, (512, 384), path="inference"))
x = torch.rand(3, 384, 512, device=dev)
res["enhance_thermal_contrast_1img_us"] = timeit(lambda: pp.enhance_thermal_contrast(x))
x1 = x[:1].repeat(3, 1, 1).contiguous()
res["enhance_thermal_contrast_1img_replicated_us"] = timeit(lambda: pp.enhance_thermal_contrast(x1))
print(json.dumps(res))
from thermal3d_vision_b200 import _lib
fn = lambda: pp.preprocess_thermal_batch(raw, (512, 384), path="inference")
out = {}
for nm in ["resize_bilinear", "fpct_sample", "fpct_classify", "fpct_select", "normalize_f32", "channels_close", "set_int"]:
    _lib.profile_begin(nm, 4096)
    for _ in range(10): fn()
    torch.cuda.synchronize()
    ms, n = _lib.profile_end()
    out[nm] = round(ms / 10 * 1e3, 1)
print(json.dumps(out))
