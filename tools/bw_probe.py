"""Developer tool: raw HBM write / copy bandwidth probes (torch kernels) to judge the write-heavy kernels against."""
import torch, json
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
res = {}
for mb in (302, 1208):
    x = torch.empty(mb * 1000 * 1000 // 4, dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    s = t(lambda: x.fill_(1.0)); res[f"fill_{mb}MB_GBs"] = round(mb / 1e3 / s, 1)
    s = t(lambda: y.copy_(x)); res[f"copy_{mb}MB_GBs_rw"] = round(2 * mb / 1e3 / s, 1)
    s = t(lambda: x.sum()); res[f"read_{mb}MB_GBs"] = round(mb / 1e3 / s, 1)
print(json.dumps(res))
