"""GPU parity: Sobel thermal enhancer (ThermalDUSt3R.preprocess_thermal) vs oracle / golden."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_sobel

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_forward_matches_reference_golden(cuda_device):
    from thermal3d_vision_b200.sobel import ThermalDUSt3R
    k = np.load(os.path.join(G, "sobel_kat.npz"))
    m = ThermalDUSt3R(torch.nn.Identity()).to(cuda_device)
    x = torch.rand(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    y = m.preprocess_thermal(x.to(cuda_device))
    assert y.requires_grad and y.shape == (2, 3, 224, 224)
    ref = ref_sobel.preprocess_thermal_torch(x, torch.tensor(0.5), torch.tensor(1.0))
    torch.testing.assert_close(y.detach().cpu(), ref, rtol=1e-5, atol=1e-6)
    assert y.double().sum().item() == pytest.approx(float(k["big_sum"]), rel=1e-6)     # 267724.249


@pytest.mark.parametrize("shape", [(2, 1, 17, 23), (1, 3, 33, 65), (3, 3, 8, 8)])
def test_forward_backward_small(cuda_device, shape):
    from thermal3d_vision_b200.sobel import sobel_enhance
    g = torch.Generator().manual_seed(shape[2])
    x = torch.rand(*shape, generator=g)
    w = torch.rand(shape[0], 3, shape[2], shape[3], generator=g)
    ew, ts = torch.tensor(0.8, requires_grad=True), torch.tensor(0.9, requires_grad=True)
    ref = ref_sobel.preprocess_thermal_torch(x, ew, ts)
    (ref * w).sum().backward()
    ew2 = torch.tensor(0.8, device=cuda_device, requires_grad=True)
    ts2 = torch.tensor(0.9, device=cuda_device, requires_grad=True)
    y = sobel_enhance(x.to(cuda_device), ew2, ts2)
    (y * w.to(cuda_device)).sum().backward()
    torch.testing.assert_close(y.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-6)
    assert ew2.grad.item() == pytest.approx(ew.grad.item(), rel=1e-4)
    assert ts2.grad.item() == pytest.approx(ts.grad.item(), rel=1e-4)


def test_golden_small_case_and_wrapper_forward(cuda_device):
    from thermal3d_vision_b200.sobel import ThermalDUSt3R
    k = np.load(os.path.join(G, "sobel_kat.npz"))
    m = ThermalDUSt3R(torch.nn.Identity()).to(cuda_device)
    with torch.no_grad():
        m.edge_weight.fill_(0.8); m.temp_scale.fill_(0.9)
    x = torch.from_numpy(k["small_in"]).to(cuda_device)
    y = m.preprocess_thermal(x)
    np.testing.assert_allclose(y.detach().cpu().numpy(), k["small_out"], rtol=1e-5, atol=1e-6)
    (y * torch.from_numpy(k["small_w"]).to(cuda_device)).sum().backward()
    assert m.edge_weight.grad.item() == pytest.approx(float(k["small_dew"]), rel=1e-4)
    assert m.temp_scale.grad.item() == pytest.approx(float(k["small_dts"]), rel=1e-4)
    # dict-style forward as DUSt3R calls it (thermal_dustr_model.py:144-156)
    class Echo(torch.nn.Module):
        def forward(self, a, b):
            return a["img"], b["img"]
    mm = ThermalDUSt3R(Echo()).to(cuda_device)
    a, b = mm({"img": x, "instance": []}, {"img": x, "instance": []})
    assert a.shape == (2, 3, 17, 23) and torch.equal(a, b)
    assert set(mm.state_dict()) == {"sobel_x", "sobel_y", "edge_weight", "temp_scale"}
