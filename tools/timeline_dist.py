"""Developer tool: tools/timeline.py under torchrun (distributed=True; argv[1] = plain|pipelined, default pipelined); rank 0 prints
the un-instrumented step time and the kernel timeline of three steps (all streams)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
from thermal3d_vision_b200 import _lib
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, rank, dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "pipelined"
step = HotPathStep(B, H, W, device=dev, distributed=True, pipelined=(mode == "pipelined"))
args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
for _ in range(10): step.run_device(*args)
step.finish(); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): step.run_device(*args)
step.finish(); e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"# mode={mode} world={dist.get_world_size()} exchange={step.exchange}  un-instrumented: {e0.elapsed_time(e1) * 10:.1f} us per step")
dist.barrier()
_lib.profile_begin("", 4096)
for _ in range(3): step.run_device(*args)
step.finish(); torch.cuda.synchronize()
tl = _lib.profile_timeline()
_lib.profile_end()
if rank == 0:
    for nm, a, b in sorted(tl, key=lambda r: r[1]):
        print(f"{a*1e3:9.1f} {b*1e3:9.1f} {(b-a)*1e3:7.1f}  {nm}")
dist.barrier(); dist.destroy_process_group()
