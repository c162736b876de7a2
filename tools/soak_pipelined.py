"""Developer tool: soak test of the pipelined step (sampling ahead, alternating output sets / streams): two different
input batches alternate for N steps; every step's packed result, and periodically its gradients / thermal planes /
metrics, must equal bit for bit what a plain (non-pipelined) step computes for the same inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from thermal3d_vision_b200.pipeline import HotPathStep
dev = torch.device("cuda:0")
B, H, W = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (64, 384, 512)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sets = [bench.make_inputs_torch(B, H, W, seed, dev) for seed in (0, 7)]
args = [tuple(d[k] for k in bench.KEYS) for d in sets]
plain = HotPathStep(B, H, W, device=dev, pipelined=False)
want = []
for a in args:
    r = plain.run_device(*a).clone()
    torch.cuda.synchronize()
    want.append({"r": r, "dpred1": plain.loss_out["dpred1"].clone(), "dconf2": plain.loss_out["dconf2"].clone(),
                 "thermal": plain.pre_both["thermal"].clone(), "metrics": plain.met_out["metrics_f64"].clone()})
step = HotPathStep(B, H, W, device=dev, pipelined=True)
bad = 0
pend = []
for k in range(N):
    j = (k * 7 // 3) & 1 if k % 5 else k & 1          # an irregular A / B pattern
    r = step.run_device(*args[j])
    pend.append((k, j, r, step.loss_out, step.pre_both, step.met_out))
    if len(pend) == 2:                                 # check the step before the one just enqueued (its buffers are still its own)
        kk, jj, rr, lo, pre, met = pend.pop(0)
        step.wait_result(rr)
        torch.cuda.current_stream().synchronize()
        ok = torch.equal(rr.view(torch.int64), want[jj]["r"].view(torch.int64))
        if kk % 10 == 0:
            ok = ok and torch.equal(lo["dpred1"], want[jj]["dpred1"]) and torch.equal(lo["dconf2"], want[jj]["dconf2"]) \
                and torch.equal(pre["thermal"], want[jj]["thermal"]) \
                and torch.equal(met["metrics_f64"].view(torch.int64), want[jj]["metrics"].view(torch.int64))
        if not ok:
            bad += 1
            print("MISMATCH at step", kk, "inputs", jj)
step.finish()
# second pass without any host synchronisation inside the loop: the host runs two steps ahead of the device, the
# pipeline runs at full speed; each step's packed result is snapshotted on the stream (ordered after the step only)
snaps = []
for k in range(N):
    j = (k * 5 // 2) & 1 if k % 3 else k & 1
    r = step.run_device(*args[j])
    if k >= 1:                                          # snapshot the PREVIOUS step (this call's lazy join ordered the stream after it)
        snaps.append((k - 1, pj, pr.clone(), plo["dpred2"][:, ::37, ::41].clone(), ppre["thermal"][:, :, ::29, ::31].clone()))
    pj, pr, plo, ppre = j, r, step.loss_out, step.pre_both
step.finish()
torch.cuda.synchronize()
plain_d2 = [w_["dpred1"] for w_ in want]
for kk, jj, rr, d2, th in snaps:
    ok = torch.equal(rr.view(torch.int64), want[jj]["r"].view(torch.int64)) and torch.equal(th, want[jj]["thermal"][:, :, ::29, ::31])
    if not ok:
        bad += 1
        print("MISMATCH (async pass) at step", kk, "inputs", jj)
print(f"soak: {N} + {N} pipelined steps, {bad} mismatches")
sys.exit(1 if bad else 0)
