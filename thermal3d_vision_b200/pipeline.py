"""The hot path as one step: thermal preprocessing -> fused thermal-aware loss
(forward + backward) -> pointmap->depth + depth metrics, batched over B pairs.

This is the public API `bench.py` times.  One "pair" = one training sample =
two 16-bit thermal frames, two predicted and two pseudo-GT pointmaps with
confidences (train_thermal_dustr.py:136-360) + the depth metrics of view 1
(utils/metrics.py:72-138).  All arithmetic runs in libt3d_sm100.so; this class
only owns buffers, streams and the (optional) data-parallel all-reduce.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from . import distributed as _dist
from . import loss as _loss
from . import metrics as _metrics
from . import preprocessing as _pre

RESULT_SIZE = 16   # packed result vector (float64): see HotPathStep.run_device


class HotPathStep:
    def __init__(self, B: int, H: int, W: int, raw_hw=(512, 640), device=None, multi_scale: bool = False,
                 alpha=0.2, edge_weight=0.5, smoothness_weight=0.3, detail_weight=0.4, distributed: bool = False,
                 exchange: str = "peer"):
        self.B, self.H, self.W, self.raw_hw = B, H, W, tuple(raw_hw)
        self.device = torch.device(device if device is not None else "cuda")
        self.kw = dict(alpha=alpha, edge_weight=edge_weight, smoothness_weight=smoothness_weight,
                       detail_weight=detail_weight, multi_scale=multi_scale)
        self.distributed = distributed
        dev = self.device
        lib = _lib.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        # outputs / workspaces are allocated once (the library never allocates)
        self.loss_out = {
            "workspace": torch.empty(lib.t3d_loss_workspace_bytes(B, H, W, 1), dtype=torch.uint8, device=dev),
            "per_sample": torch.empty(B, 8, **f32), "batch": torch.empty(8, **f32),
            "dpred1": torch.empty(B, H, W, 3, **f32), "dpred2": torch.empty(B, H, W, 3, **f32),
            "dconf1": torch.empty(B, H, W, **f32), "dconf2": torch.empty(B, H, W, **f32),
        }
        # both views are preprocessed by ONE call when the caller hands them over as the two halves of a single
        # [2B,...] tensor (what run_host's staging does): outputs are halves of one buffer as well
        self.pre_both = {
            "thermal": torch.empty(2 * B, 3, H, W, **f32),
            "percentiles": torch.empty(2 * B, 2, dtype=torch.float64, device=dev),
            "histogram": torch.empty(2 * B, 65536, dtype=torch.int32, device=dev),
            "grad_stats": torch.empty(2 * B, max(lib.t3d_preprocess_stats_tiles(H, W), 1), 4, **f32),
            "workspace": torch.empty(lib.t3d_preprocess_workspace_bytes(2 * B, H, W), dtype=torch.uint8, device=dev),
        }
        self.pre_out = [{k: (v[:B] if i == 0 else v[B:]) if k != "workspace" else
                         torch.empty(lib.t3d_preprocess_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
                         for k, v in self.pre_both.items()} for i in range(2)]
        self.met_out = {
            "workspace": torch.empty(lib.t3d_depth_metrics_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev),
            "metrics": torch.empty(B, 8, **f32), "metrics_f64": torch.empty(B, 8, dtype=torch.float64, device=dev),
            "medians": torch.empty(B, 2, **f32),
        }
        # side streams: the two preprocessing calls and the metric pipeline are independent of each other
        # (many short, latency-bound kernels) and overlap; the loss needs both thermal batches and then
        # runs alone on the caller's stream
        # (high priority: the metric pipeline's small one-CTA-per-image kernels then get the next free SM slots
        # instead of queueing behind the preprocessing kernels' remaining CTAs -- 0.465 -> 0.455 ms per step)
        self.side = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(2)]
        self.fork = torch.cuda.Event()
        self.joins = [torch.cuda.Event() for _ in range(2)]
        self.overlap = True
        self._shared_hint = None
        # percentiles from sampled value windows (bit-identical to the exact-histogram path, no per-pixel atomic);
        # set True to also get the 65 536-bin histograms of the resized frames in pre_both["histogram"]
        self.histogram = False
        # two packed-result vectors used alternately: with distributed=True the all-reduce of step i is asynchronous
        # and overlaps the kernels of step i + 1 (it is a 128-byte, latency-bound collective)
        self.results = [torch.zeros(RESULT_SIZE, dtype=torch.float64, device=dev) for _ in range(2)]
        self.pending = [None, None]
        self.calls = 0
        self.result = self.results[0]
        self.result_host = torch.zeros(RESULT_SIZE, dtype=torch.float64).pin_memory()
        self.staging = None        # two device staging sets of run_host (allocated on first use)
        # Data-parallel exchange of the packed result.  "peer" (NCCL process group on one NVLink node): every rank
        # owns a 4 KB mailbox in symmetric memory; the step's epilogue kernel stores the rank's 16 doubles straight
        # into every peer's mailbox and a one-warp kernel adds them up one step later -- no collective call and no
        # host work per step beyond two launches (a per-step dist.all_reduce costs enough host time to make the
        # 0.43 ms step host-bound: 0.458 vs 0.435 ms).  "nccl": one asynchronous dist.all_reduce per step.
        self.exchange = None
        if distributed and _dist.world()[1] > 1:
            import os
            exchange = os.environ.get("T3D_EXCHANGE", exchange)      # tuning knob
            if exchange not in ("peer", "nccl"):
                raise ValueError("exchange must be 'peer' or 'nccl'")
            self.exchange = exchange
        if self.exchange == "peer":
            import ctypes as C
            import warnings
            import torch.distributed as dist
            rank, world = _dist.world()
            if world > 16:
                raise ValueError("peer exchange: at most 16 ranks on one node (include/t3d.h T3D_MAX_PEERS)")
            nbytes = int(lib.t3d_mailbox_bytes())
            err = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.mailbox = symm_mem.empty(nbytes // 8, dtype=torch.float64, device=dev)
                self.mailbox.zero_()
                hdl = symm_mem.rendezvous(self.mailbox, dist.group.WORLD.group_name)
                self.peer_ptrs = (C.c_uint64 * world)(*[int(p) for p in hdl.buffer_ptrs])
                self._symm_handle = hdl
            except Exception as e:      # no peer-addressable memory on this system (e.g. GPUs without P2P access)
                err = e
            # the ranks must agree: one that cannot map its peers takes everybody to the NCCL exchange (still on the GPUs)
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                warnings.warn(f"peer-memory exchange unavailable ({err!r}); using one NCCL all-reduce per step")
                self.exchange = "nccl"
            else:
                self.local = [torch.zeros(RESULT_SIZE, dtype=torch.float64, device=dev) for _ in range(2)]
                self.reduced = [True, True]     # results[i] holds the global vector of the step that last used slot i
                self.step_of = [-1, -1]
                torch.cuda.synchronize(dev)
                dist.barrier()                  # every mailbox is zeroed before anybody's first peer store

    # ------------------------------------------------------------------ bytes
    def algorithmic_bytes(self) -> Dict[str, int]:
        B, n = self.B, self.H * self.W
        raw = self.raw_hw[0] * self.raw_hw[1]
        return {
            # read 2x12 pred, 2x12 gt, 2x4 conf, 2x4 thermal (the 3 planes are replicas: one is read), write 32
            "loss": 96 * B * n,
            "preprocess": 2 * B * (2 * raw + 12 * n),          # u16 frame in, 3 fp32 planes out
            "metrics": B * (12 * n + 4 * n + 64),              # pointmap (AoS sectors) + GT in, 64 B out
        }

    # ------------------------------------------------------------------ device-resident step
    def run_device(self, raw1, raw2, pred1, pred2, gt1, gt2, conf1, conf2, gt_depth):
        """All inputs already in HBM.  Returns the packed device result vector (float64):
        [0] sum of valid per-sample losses  [1..4] sums of components  [5] n_valid  [6] B
        [7..13] sums of finite per-image metrics (abs_rel..acc_3)  [14] n_images  [15] unused.
        With distributed=True the vector is summed over the ranks (peer-memory mailboxes, or one NCCL all-reduce with
        exchange="nccl"); it holds the global sums once wait_result() / finish() / the next call has been enqueued."""
        size = (self.W, self.H)
        B = self.B
        main = torch.cuda.current_stream(self.device)
        stacked = (raw1.is_contiguous() and raw2.is_contiguous() and raw1.shape == raw2.shape and
                   raw1.untyped_storage().data_ptr() == raw2.untyped_storage().data_ptr() and
                   raw2.storage_offset() == raw1.storage_offset() + raw1.numel())      # halves of one tensor
        if stacked:
            raw_both = torch.as_strided(raw1, (2 * B,) + tuple(raw1.shape[1:]), raw1.stride(), raw1.storage_offset())

        def preprocess_all():
            if stacked:
                tb = _pre.preprocess_thermal_batch(raw_both, size, path="train", out=self.pre_both, histogram=self.histogram)
                gs = tb.grad_stats
                return (tb.thermal[:B], tb.thermal[B:]), (None, None) if gs is None else (gs[:B], gs[B:])
            a = _pre.preprocess_thermal_batch(raw1, size, path="train", out=self.pre_out[0], histogram=self.histogram)
            b = _pre.preprocess_thermal_batch(raw2, size, path="train", out=self.pre_out[1], histogram=self.histogram)
            return (a.thermal, b.thermal), (a.grad_stats, b.grad_stats)

        if self.overlap != self._shared_hint:       # the metric pipeline shares the SMs with the preprocessing
            self._shared_hint = self.overlap
            _lib.lib().t3d_preprocess_set_shared(1 if self.overlap else 0)
        if self.overlap:
            self.fork.record(main)
            with torch.cuda.stream(self.side[1]):           # depth metrics (Z of pred1 read in place)
                self.side[1].wait_event(self.fork)
                me = _metrics.compute_depth_metrics_batch(pred1, gt_depth, out=self.met_out)
                self.joins[1].record(self.side[1])
            (t1, t2), stats = preprocess_all()
            main.wait_event(self.joins[1])
        else:
            (t1, t2), stats = preprocess_all()
            me = _metrics.compute_depth_metrics_batch(pred1, gt_depth, out=self.met_out)
        # the normalisation kernel already summed the thermal gradients: the loss skips its statistics pass
        lo = _loss.fused_thermal_loss_fwd_bwd(pred1, pred2, gt1, gt2, conf1, conf2, t1, t2,
                                              out=self.loss_out, thermal_stats=stats,
                                              thermal_replicated=True,   # preprocess_thermal_batch wrote 3 identical planes
                                              grad_scale=_dist.global_grad_scale(self.B) if self.distributed else None,
                                              rescale_invalid=False,     # done by t3d_step_epilogue below
                                              **self.kw)
        i = self.calls & 1
        step_no = self.calls
        self.calls += 1
        lib = _lib.lib()
        stream = _lib.current_stream_ptr()
        if self.exchange == "peer":
            # reduce(step - 1) is enqueued before epilogue(step): the order that makes the two mailbox slots reusable
            self._reduce_pending()
            rank, world = _dist.world()
            rc = lib.t3d_step_epilogue_peers(_lib.ptr(lo["dpred1"]), _lib.ptr(lo["dpred2"]), _lib.ptr(lo["dconf1"]),
                                             _lib.ptr(lo["dconf2"]), _lib.ptr(lo["per_sample"]), _lib.ptr(lo["batch"]),
                                             _lib.ptr(me["metrics_f64"]), self.B, self.H, self.W, self.B,
                                             _lib.ptr(self.local[i]), self.peer_ptrs, world, rank, step_no, stream)
            _lib.check(rc, "t3d_step_epilogue_peers")
            self.reduced[i] = False
            self.step_of[i] = step_no
            self.result = self.results[i]          # global once wait_result() / the next step has enqueued the reduction
            return self.result
        if self.pending[i] is not None:
            self.pending[i].wait()                 # the all-reduce of step - 2 is done with this vector
            self.pending[i] = None
        r = self.results[i]
        # validity fix-up of the gradients + packing of the step's scalars: one launch
        rc = lib.t3d_step_epilogue(_lib.ptr(lo["dpred1"]), _lib.ptr(lo["dpred2"]), _lib.ptr(lo["dconf1"]),
                                   _lib.ptr(lo["dconf2"]), _lib.ptr(lo["per_sample"]), _lib.ptr(lo["batch"]),
                                   _lib.ptr(me["metrics_f64"]), self.B, self.H, self.W, self.B,
                                   _lib.ptr(r), stream)
        _lib.check(rc, "t3d_step_epilogue")
        if self.exchange == "nccl":
            self.pending[i] = _dist.all_reduce_result(r, async_op=True)
        self.result = r
        return r

    def capture_graph(self, raw1, raw2, pred1, pred2, gt1, gt2, conf1, conf2, gt_depth):
        """Capture one `run_device` on these (static) device tensors into a CUDA graph and return `replay()`, which
        re-runs the whole step (both streams, ~15 kernels) with one launch and returns the packed result vector.
        For launch-bound shapes (BASELINE configs[1]: batch 8 at 224x224 is ~45 MB of traffic, a few microseconds
        of HBM time) this is what removes the per-kernel launch latency.  Single-process only (the all-reduce of a
        distributed step is issued outside any graph)."""
        if self.distributed:
            raise ValueError("capture_graph is for single-process steps")
        args = (raw1, raw2, pred1, pred2, gt1, gt2, conf1, conf2, gt_depth)
        cur = torch.cuda.current_stream(self.device)
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            for _ in range(2):                     # one-time attribute / occupancy queries happen outside the capture
                self.run_device(*args)
        cur.wait_stream(warm)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            r = self.run_device(*args)

        def replay():
            graph.replay()
            return r
        replay.graph = graph
        return replay

    def _reduce_pending(self):
        """peer exchange: enqueue the reduction of every step whose global vector has not been formed yet (oldest first)."""
        if self.exchange != "peer":
            return
        lib, stream, world = _lib.lib(), _lib.current_stream_ptr(), _dist.world()[1]
        for i in sorted(range(2), key=lambda k: self.step_of[k]):
            if not self.reduced[i]:
                rc = lib.t3d_mailbox_reduce(_lib.ptr(self.mailbox), world, self.step_of[i], _lib.ptr(self.results[i]), stream)
                _lib.check(rc, "t3d_mailbox_reduce")
                self.reduced[i] = True

    def wait_result(self, r: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Order the current stream after the (asynchronous) all-reduce of `r` (default: the latest result).
        Call before reading a result of a distributed step; a no-op otherwise."""
        r = self.result if r is None else r
        self._reduce_pending()
        for i in range(2):
            if self.results[i] is r and self.pending[i] is not None:
                self.pending[i].wait()
                self.pending[i] = None
        return r

    def finish(self):
        """Order the current stream after every outstanding all-reduce (end of a run / of a timed region)."""
        self._reduce_pending()
        for i in range(2):
            if self.pending[i] is not None:
                self.pending[i].wait()
                self.pending[i] = None

    # ------------------------------------------------------------------ host-buffer step (e2e)
    def run_host(self, host: Dict[str, torch.Tensor]):
        """Inputs are PINNED HOST tensors (raw1, raw2 uint16; pred1, pred2, gt1, gt2, conf1, conf2, gt_depth float32).
        Copies them to the device, runs the step, copies the packed result back to pinned host memory.
        Asynchronous; `result_host` is valid after a synchronise of the current stream.

        The H2D copies go through a dedicated copy stream into one of two staging sets, so the copies of call
        i + 1 overlap the kernels of call i (the step is PCIe-bound: 839 MB of inputs per call at batch 64);
        every call still copies all of its own inputs and reads its own result back."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if self.staging is None:
            self.staging = []
            for _ in range(2):
                st = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items() if k not in ("raw1", "raw2")}
                both = torch.empty((2 * self.B,) + tuple(host["raw1"].shape[1:]), dtype=host["raw1"].dtype, device=dev)
                st["raw1"], st["raw2"] = both[:self.B], both[self.B:]
                self.staging.append(st)
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.copied = [torch.cuda.Event() for _ in range(2)]
            self.consumed = [torch.cuda.Event() for _ in range(2)]
            self.host_calls = 0
        i = self.host_calls & 1
        s = self.staging[i]
        with torch.cuda.stream(self.copy_stream):
            if self.host_calls >= 2:
                self.copy_stream.wait_event(self.consumed[i])      # the kernels of call - 2 are done with this set
            else:
                self.copy_stream.wait_stream(main)
            for k, v in host.items():
                s[k].copy_(v, non_blocking=True)
            self.copied[i].record(self.copy_stream)
        main.wait_event(self.copied[i])
        r = self.run_device(s["raw1"], s["raw2"], s["pred1"], s["pred2"], s["gt1"], s["gt2"], s["conf1"], s["conf2"],
                            s["gt_depth"])
        self.consumed[i].record(main)
        self.wait_result(r)
        self.result_host.copy_(r, non_blocking=True)
        self.host_calls += 1
        return self.result_host

    @staticmethod
    def h2d_bytes(host: Dict[str, torch.Tensor]) -> int:
        return int(sum(v.numel() * v.element_size() for v in host.values()))

    @staticmethod
    def summarize(result_host: torch.Tensor) -> Dict[str, float]:
        return _dist.summarize(result_host.tolist())


class EvalStep:
    """BASELINE.json configs[4]: full-resolution 16-bit frames -> preprocessing (the model's input) and, for the
    model's pointmaps (the caller's; synthetic in bench.py), pointmap -> depth -> depth metrics against the GT
    depth, accumulated over a dataset shard with the reference's semantics (utils/metrics.py:86-136: finite
    per-image metrics summed, divided by the count of ALL images).  One rank = one contiguous shard
    (distributed.shard_range); `finish()` does the single all-reduce of the 8-double accumulator.

    path='train': data/dataset_loader.py:237-249 + enhance_thermal_contrast (u16 resize, raw counts);
    path='inference': utils/evaluate_depth_metrics.py:162-197 (/65535, float resize, float percentiles).
    GT depth of another size is nearest-resampled inside the metric kernels (utils/evaluate_depth_metrics.py:320-323).
    """

    def __init__(self, B: int, H: int, W: int, raw_hw=(512, 640), gt_hw=None, device=None, path: str = "train"):
        self.B, self.H, self.W, self.raw_hw, self.path = B, H, W, tuple(raw_hw), path
        self.gt_hw = tuple(gt_hw) if gt_hw is not None else (H, W)
        self.device = torch.device(device if device is not None else "cuda")
        dev, lib = self.device, _lib.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        self.pre_out = {"thermal": torch.empty(B, 3, H, W, **f32),
                        "percentiles": torch.empty(B, 2, dtype=torch.float64, device=dev),
                        "workspace": torch.empty(lib.t3d_preprocess_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)}
        self.met_out = {"workspace": torch.empty(lib.t3d_depth_metrics_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev),
                        "metrics": torch.empty(B, 8, **f32), "metrics_f64": torch.empty(B, 8, dtype=torch.float64, device=dev),
                        "medians": torch.empty(B, 2, **f32)}
        self.acc = _metrics.MetricAccumulator(dev)
        self.side = torch.cuda.Stream(device=dev, priority=-1)
        self.fork, self.join = torch.cuda.Event(), torch.cuda.Event()
        lib.t3d_preprocess_set_shared(1)            # the metric pipeline runs beside the preprocessing (run_batch)

    def algorithmic_bytes(self) -> int:
        """Per batch: u16 frame in + 3 fp32 planes out, AoS pointmap + GT depth in."""
        n, raw, gt = self.H * self.W, self.raw_hw[0] * self.raw_hw[1], self.gt_hw[0] * self.gt_hw[1]
        return self.B * ((2 * raw + 12 * n) + (12 * n + 4 * min(gt, n) + 64))

    def run_batch(self, raw, pointmap, gt_depth):
        """raw [B,Hs,Ws] uint16, pointmap [B,H,W,3] float32, gt_depth [B,gh,gw] float32, all on the device.
        Returns the preprocessed thermal batch [B,3,H,W]; the metrics go into the accumulator.  No host sync."""
        main = torch.cuda.current_stream(self.device)
        self.fork.record(main)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.fork)
            me = _metrics.compute_depth_metrics_batch(pointmap, gt_depth, out=self.met_out)
            self.join.record(self.side)
        tb = _pre.preprocess_thermal_batch(raw, (self.W, self.H), path=self.path, out=self.pre_out, histogram=False)
        main.wait_event(self.join)
        self.acc.update(me["metrics_f64"])
        return tb.thermal

    def finish(self) -> Dict[str, float]:
        """All-reduce (SUM) the accumulator over the ranks and return the dataset-mean metrics (host sync)."""
        return self.acc.all_reduce().result()
