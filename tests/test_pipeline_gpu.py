"""GPU parity of the benchmarked path itself: pipeline.HotPathStep (preprocessing of both views -> fused loss
forward + backward -> pointmap -> depth -> metrics -> packed result) against the oracle loop that mirrors
train_thermal_dustr.py:182-360 + utils/metrics.py:72-138 on the same inputs."""
import numpy as np
import pytest
import torch

from oracle import ref_loss, ref_metrics, ref_preprocess

pytestmark = pytest.mark.gpu
KW = dict(alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4)


def _oracle_step(raw1, raw2, P1, P2, G1, G2, C1, C2, gt_depth, H, W, multi):
    B = P1.shape[0]
    T1 = torch.from_numpy(np.stack([ref_preprocess.train_path(raw1[i], (H, W))[0] for i in range(B)]))
    T2 = torch.from_numpy(np.stack([ref_preprocess.train_path(raw2[i], (H, W))[0] for i in range(B)]))
    lead = [x.clone().requires_grad_() for x in (P1, P2, C1, C2)]
    mean, rows, valid = ref_loss.batched_loss_torch(lead[0], lead[1], G1, G2, lead[2], lead[3], T1, T2,
                                                    multi_scale=multi, **KW)
    mean.backward()
    per = [ref_metrics.compute_depth_metrics(P1[i, ..., 2].numpy(), gt_depth[i].numpy()) for i in range(B)]
    return T1, T2, mean.item(), rows, valid, [x.grad for x in lead], ref_metrics.accumulate_dataset(per)


@pytest.mark.parametrize("B,H,W", [(3, 96, 160), (1, 224, 224), (2, 100, 260)])
@pytest.mark.parametrize("multi", [False, True])
def test_hot_path_step_matches_oracle_loop(cuda_device, multi, B, H, W):
    from thermal3d_vision_b200.pipeline import HotPathStep
    raw = ref_preprocess.make_raw_frames(2 * B, seed=13, hw=(128, 200))
    P1, P2, G1, G2, C1, C2, _, _ = ref_loss.make_batch_inputs(B, H, W, seed=17, stress_conf=True)
    gt_depth = G1[..., 2].clone()
    gt_depth[B - 1, :7] = 0.0                               # invalid GT region in one image
    T1, T2, mean, rows, valid, grads, metrics = _oracle_step(raw[:B], raw[B:], P1, P2, G1, G2, C1, C2, gt_depth, H, W, multi)

    step = HotPathStep(B, H, W, raw_hw=(128, 200), device=cuda_device, multi_scale=multi, **KW)
    both = torch.from_numpy(raw).to(cuda_device)
    d = [x.to(cuda_device) for x in (P1, P2, G1, G2, C1, C2, gt_depth)]
    r = step.run_device(both[:B], both[B:], *d).cpu()
    s = HotPathStep.summarize(r)
    # the thermal batches fed to the loss are bit-exact, hence everything downstream is comparable
    assert np.array_equal(step.pre_both["thermal"][:B].cpu().numpy(), T1.numpy())
    assert np.array_equal(step.pre_both["thermal"][B:].cpu().numpy(), T2.numpy())
    assert s["n_valid"] == float(valid.sum()) and s["n_pairs"] == B and s["n_images"] == B
    assert s["loss"] == pytest.approx(mean, rel=1e-5)
    np.testing.assert_allclose(step.loss_out["per_sample"][:, :5].cpu().numpy(), rows, rtol=1e-5)
    for name, got, ref in (("dpred1", step.loss_out["dpred1"], grads[0]), ("dpred2", step.loss_out["dpred2"], grads[1]),
                           ("dconf1", step.loss_out["dconf1"], grads[2]), ("dconf2", step.loss_out["dconf2"], grads[3])):
        torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-6, msg=lambda m: f"{name}: {m}")
    for k in ref_metrics.KEYS7:
        assert s[k] == pytest.approx(metrics[k], rel=1e-5), k
    # the host-buffer entry point (pinned H2D inside) gives the same packed vector, bit for bit
    host = {"raw1": torch.from_numpy(raw[:B]).pin_memory(), "raw2": torch.from_numpy(raw[B:]).pin_memory(),
            "pred1": P1.pin_memory(), "pred2": P2.pin_memory(), "gt1": G1.pin_memory(), "gt2": G2.pin_memory(),
            "conf1": C1.pin_memory(), "conf2": C2.pin_memory(), "gt_depth": gt_depth.pin_memory()}
    rh = step.run_host(host)
    torch.cuda.synchronize()
    assert torch.equal(rh, r)


def test_hot_path_step_full_size_is_deterministic(cuda_device):
    """BASELINE configs[2] shape (512x384) at a small batch: two runs give bit-identical results and gradients;
    the histogram option does not change a bit."""
    from thermal3d_vision_b200.pipeline import HotPathStep
    import bench
    B, H, W = 4, 384, 512
    d = bench.make_inputs_torch(B, H, W, seed=5, device=cuda_device)
    args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
    step = HotPathStep(B, H, W, device=cuda_device)
    a = step.run_device(*args).clone()
    ga = [step.loss_out[k].clone() for k in ("dpred1", "dpred2", "dconf1", "dconf2")]
    b = step.run_device(*args).clone()
    assert torch.equal(a, b)
    for x, k in zip(ga, ("dpred1", "dpred2", "dconf1", "dconf2")):
        assert torch.equal(x, step.loss_out[k]), k
    step.histogram = True
    c = step.run_device(*args).clone()
    assert torch.equal(a, c)
    assert (step.pre_both["histogram"].sum(1) == H * W).all()
    s = HotPathStep.summarize(a.cpu())
    assert s["n_valid"] == B and 0 < s["loss"] < 100 and 0 <= s["abs_rel"] < 1


def test_graph_replay_equals_stream_launches(cuda_device):
    """HotPathStep.capture_graph: the whole step as one CUDA-graph launch gives the bits of the eager step."""
    from thermal3d_vision_b200.pipeline import HotPathStep
    import bench
    B, H, W = 8, 224, 224                               # BASELINE configs[1] shape: launch-bound
    d = bench.make_inputs_torch(B, H, W, seed=2, device=cuda_device)
    args = (d["raw1"], d["raw2"], d["pred1"], d["pred2"], d["gt1"], d["gt2"], d["conf1"], d["conf2"], d["gt_depth"])
    step = HotPathStep(B, H, W, device=cuda_device)
    eager = step.run_device(*args).clone()
    grads = [step.loss_out[k].clone() for k in ("dpred1", "dconf2")]
    replay = step.capture_graph(*args)
    for _ in range(3):
        for k in ("dpred1", "dconf2"):
            step.loss_out[k].zero_()
        r = replay()
        torch.cuda.synchronize()
        assert torch.equal(r, eager)
        assert torch.equal(step.loss_out["dpred1"], grads[0]) and torch.equal(step.loss_out["dconf2"], grads[1])


def test_hot_path_step_invalid_sample(cuda_device):
    """A sample whose loss is not finite is skipped as train_thermal_dustr.py:320 does: it drops out of the mean,
    its gradients are exact zeros and the other samples' gradients carry 1 / n_valid (loss_finalize_kernel +
    t3d_step_epilogue do this on the device, without a host round trip)."""
    from thermal3d_vision_b200.pipeline import HotPathStep
    B, H, W = 3, 64, 128
    raw = ref_preprocess.make_raw_frames(2 * B, seed=3, hw=(96, 160))
    P1, P2, G1, G2, C1, C2, _, _ = ref_loss.make_batch_inputs(B, H, W, seed=23)
    P1[1, 5, 7, 2] = float("nan")
    gt_depth = G1[..., 2].clone()
    T1, T2, mean, rows, valid, grads, metrics = _oracle_step(raw[:B], raw[B:], P1, P2, G1, G2, C1, C2, gt_depth, H, W, False)
    assert valid.tolist() == [True, False, True]
    step = HotPathStep(B, H, W, raw_hw=(96, 160), device=cuda_device, **KW)
    both = torch.from_numpy(raw).to(cuda_device)
    d = [x.to(cuda_device) for x in (P1, P2, G1, G2, C1, C2, gt_depth)]
    for _ in range(2):                                      # second call: the workspace counters were reset
        s = HotPathStep.summarize(step.run_device(both[:B], both[B:], *d).cpu())
        assert s["n_valid"] == 2.0 and s["n_pairs"] == B
        assert s["loss"] == pytest.approx(mean, rel=1e-5)
        for name, got, ref in (("dpred1", step.loss_out["dpred1"], grads[0]), ("dpred2", step.loss_out["dpred2"], grads[1]),
                               ("dconf1", step.loss_out["dconf1"], grads[2]), ("dconf2", step.loss_out["dconf2"], grads[3])):
            ref = torch.zeros_like(got.cpu()) if ref is None else ref
            assert torch.count_nonzero(got[1]) == 0, name
            torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-6, msg=lambda m: f"{name}: {m}")


def test_peer_mailbox_exchange_world_of_one(cuda_device):
    """t3d_step_epilogue_peers + t3d_mailbox_reduce (the data-parallel exchange over peer memory) with this process as
    the only rank: the reduced vector is the packed vector of t3d_step_epilogue, step after step (both parities of
    the mailbox get reused); with two "ranks" pointing at the same mailbox the slots add up in rank order."""
    import ctypes as C
    from thermal3d_vision_b200 import _lib
    lib = _lib.lib()
    B, H, W = 3, 8, 16
    g = torch.Generator().manual_seed(5)
    per_sample = torch.rand(B, 8, generator=g).to(cuda_device)
    per_sample[:, 5] = torch.tensor([1.0, 0.0, 1.0])                         # sample 1 invalid
    batch = torch.tensor([0, 0, 0, 0, 0, 2.0, float(B), 0], dtype=torch.float32, device=cuda_device)
    m64 = torch.rand(B, 8, generator=g, dtype=torch.float64).to(cuda_device)
    grads = [torch.ones(B, H, W, 3, device=cuda_device), torch.ones(B, H, W, 3, device=cuda_device),
             torch.ones(B, H, W, device=cuda_device), torch.ones(B, H, W, device=cuda_device)]
    ref = torch.zeros(16, dtype=torch.float64, device=cuda_device)
    st = _lib.current_stream_ptr()
    _lib.check(lib.t3d_step_epilogue(*[_lib.ptr(x) for x in grads], _lib.ptr(per_sample), _lib.ptr(batch), _lib.ptr(m64),
                                     B, H, W, B, _lib.ptr(ref), st), "t3d_step_epilogue")
    assert torch.count_nonzero(grads[0][1]) == 0 and torch.all(grads[0][0] == 1.5)    # fix-up: B / n_valid = 3 / 2
    mailbox = torch.zeros(int(lib.t3d_mailbox_bytes()) // 8, dtype=torch.float64, device=cuda_device)
    batch_ok = batch.clone(); batch_ok[5] = float(B)                              # no second fix-up of the gradients
    for world in (1, 2):
        mailbox.zero_()
        peers = (C.c_uint64 * world)(*([mailbox.data_ptr()] * world))
        for step in range(5):
            local = torch.zeros(16, dtype=torch.float64, device=cuda_device)
            out = torch.full((16,), -1.0, dtype=torch.float64, device=cuda_device)
            for rank in range(world):                                             # every "rank" posts into the one mailbox
                _lib.check(lib.t3d_step_epilogue_peers(*[_lib.ptr(x) for x in grads], _lib.ptr(per_sample), _lib.ptr(batch_ok),
                                                       _lib.ptr(m64), B, H, W, B, _lib.ptr(local), peers, world, rank, step, st),
                           "t3d_step_epilogue_peers")
            _lib.check(lib.t3d_mailbox_reduce(_lib.ptr(mailbox), world, step, _lib.ptr(out), st), "t3d_mailbox_reduce")
            assert torch.equal(local, ref)
            assert torch.equal(out, ref * world), (world, step)
