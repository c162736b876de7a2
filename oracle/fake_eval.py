"""Oracle helper (test infrastructure): a deterministic stand-in model and dataloader for
`evaluate_thermal_depth(model, dataloader, device)` (utils/metrics.py:72-138).

The DUSt3R ViT is out of scope, so the behavioural test drives both the live reference function (build container,
oracle/gen_golden.py -> tests/golden/evaluate_kat.npz) and ours (GPU box) with the same fake model: its "pointmap"
is a fixed function of the thermal input, returned in each of the output conventions the reference's loop accepts
(utils/metrics.py:105-117: a tuple whose first element is a dict with 'pts3d', a dict with 'pred1', a bare tensor;
with or without a leading batch dimension).
"""
import numpy as np
import torch


class FakeModel(torch.nn.Module):
    def __init__(self, convention: str):
        super().__init__()
        self.convention = convention
        self.calls = 0

    def forward(self, view1, view2):
        img = view1["img"]                                   # [1,3,H,W]
        assert img.shape[0] == 1 and view1["instance"] == [] and view2 is view1      # monocular call of :103-104
        self.calls += 1
        x = img[0, 0]
        H, W = x.shape
        v, u = torch.meshgrid(torch.arange(H, device=x.device, dtype=torch.float32),
                              torch.arange(W, device=x.device, dtype=torch.float32), indexing="ij")
        z = 1.0 + 5.0 * x + 0.05 * torch.sin(0.37 * u + 0.11 * v)            # a depth-like function of the input
        pm = torch.stack([(u - W / 2) * z / 300.0, (v - H / 2) * z / 300.0, z], dim=-1)       # [H,W,3]
        if self.convention == "tuple_dict_batched":
            return {"pts3d": pm.unsqueeze(0), "conf": torch.ones(1, H, W, device=x.device)}, {"pts3d": pm.unsqueeze(0)}
        if self.convention == "dict_pred1":
            return {"pred1": {"pts3d": pm}, "pred2": {"pts3d": pm}}
        if self.convention == "tuple_tensor":
            return pm.unsqueeze(0), pm.unsqueeze(0)
        raise ValueError(self.convention)


def make_loader(seed: int = 0, H: int = 48, W: int = 64):
    """A list of batches like FreiburgDataset's collate output: thermal1 [B,3,H,W], depth1 [B,H,W] (or absent / None)."""
    g = torch.Generator().manual_seed(seed)
    batches = []
    for bsz in (3, 2, 1, 2):
        th = torch.rand(bsz, 1, H, W, generator=g).repeat(1, 3, 1, 1)
        z = 1.0 + 5.0 * th[:, 0]
        depth = z * (1.0 + 0.2 * torch.randn(bsz, H, W, generator=g)).abs() * 1.3 + 0.05
        depth[:, : H // 6] = 0.0                                             # no GT there (mask = gt > 0)
        batches.append({"thermal1": th, "depth1": depth})
    batches[1]["depth1"][1, 10, 10] = float("inf")                           # non-finite GT pixel: masked out
    batches[3]["thermal1"][0, :, 30, 30] = float("nan")                      # NaN prediction in the mask: that sample's four
    #                                                                          error metrics are NaN -> skipped, but the sample
    #                                                                          still counts in the denominator (:128-136)
    batches.insert(2, {"thermal1": torch.rand(2, 3, H, W, generator=g)})     # no depth: skipped (:96)
    batches.insert(4, {"thermal1": torch.rand(1, 3, H, W, generator=g), "depth1": None})
    return batches


KEYS = ("abs_rel", "sq_rel", "rmse", "rmse_log", "acc_1", "acc_2", "acc_3")
CONVENTIONS = ("tuple_dict_batched", "dict_pred1", "tuple_tensor")


def as_vector(result: dict) -> np.ndarray:
    return np.array([float(result[k]) for k in KEYS], np.float64)
