"""Developer tool (torchrun, N ranks): the box's aggregate pinned-host -> device rate, the ceiling of the e2e leg.

Every rank copies one step's inputs (839 MB at batch 64) from pinned host memory to its GPU, all ranks at once:
default pinned pages vs write-combined pinned pages (cudaHostAllocWriteCombined), one vs two copy streams."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NBYTES = 838860800
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")


def alloc(wc):
    if not wc:
        return torch.empty(NBYTES, dtype=torch.uint8).pin_memory(), None
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NBYTES), ctypes.c_uint(0x04))      # cudaHostAllocWriteCombined
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * NBYTES).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p


def run(host, streams, iters=8):
    dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    ss = [torch.cuda.Stream(device=dev) for _ in range(streams)]
    chunk = NBYTES // streams
    def once():
        for k, s in enumerate(ss):
            with torch.cuda.stream(s):
                dst[k * chunk:(k + 1) * chunk].copy_(host[k * chunk:(k + 1) * chunk], non_blocking=True)
    for _ in range(2): once()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): once()
    for s in ss: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


res = {"world": world, "bytes_per_rank": NBYTES}
for wc in (False, True):
    try:
        host, keep = alloc(wc)
        host[::4096] = 1                               # touch the pages
    except Exception as e:
        res[f"wc={wc}"] = repr(e); continue
    for streams in (1, 2):
        ms = run(host, streams)
        res[f"wc={int(wc)}_streams={streams}"] = {"ms": round(ms, 3), "GBs_per_rank": round(NBYTES / ms / 1e6, 1),
                                                  "GBs_aggregate": round(world * NBYTES / ms / 1e6, 1)}
    del host
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
