"""Developer tool (torchrun, >= 2 GPUs): HotPathStep with the peer-memory exchange vs one NCCL all-reduce per step --
same global vectors, and the step time of both."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
B, H, W = 64, 384, 512
d = bench.make_inputs_torch(B, H, W, rank, dev)
args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
out = {}
for mode in ("nccl", "peer"):
    step = HotPathStep(B, H, W, device=dev, distributed=True, exchange=mode)
    vecs = []
    for k in range(7):
        r = step.run_device(*args)
        if k % 3 == 0:
            vecs.append(step.wait_result(r).clone())        # read some results right away, some one step later
    vecs.append(step.wait_result().clone())
    step.finish(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): step.run_device(*args)
    step.finish(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 5], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[mode] = (torch.stack(vecs).cpu(), t.item())
    dist.barrier()
a, b = out["nccl"][0], out["peer"][0]
same = torch.allclose(a, b, rtol=1e-12, atol=0)
g = [None] * dist.get_world_size()
dist.all_gather_object(g, b[-1].tolist())
if rank == 0:
    print("vectors equal (rtol 1e-12):", same, "bitwise:", torch.equal(a, b), "identical on all ranks:", all(x == g[0] for x in g))
    print("n_pairs", b[-1][6].item(), "n_valid", b[-1][5].item(), "loss", (b[-1][0] / b[-1][5]).item())
    print("step_us nccl", round(out["nccl"][1], 1), "peer", round(out["peer"][1], 1))
dist.barrier(); dist.destroy_process_group()
