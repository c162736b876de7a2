"""GPU parity of the experimental fire-scene pipeline (thermal3d_vision_b200/fire.py -> csrc/t3d_fire.cu) against
tests/golden/fire_kat.npz -- outputs of the LIVE reference functions (thermal_dustr_inference_for_experiment.py:62-377,
np.random seeded) and of the stock cv2 / numpy operators, written by oracle/gen_golden.py in the build container --
and against the NumPy restatements in oracle/ref_fire.py on further shapes.
Bar: byte / integer operators (CLAHE, Canny, histogram) bit-exact; float operators to rounding (tolerances below)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_fire

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = [(96, 128), (75, 100)]


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(G, "fire_kat.npz"))


@pytest.mark.parametrize("h,w", SIZES)
def test_operators_match_stock_libraries(cuda_device, kat, h, w):
    from thermal3d_vision_b200 import fire
    tag = f"{h}x{w}"
    u8 = torch.from_numpy(kat["u8_" + tag]).to(cuda_device)
    for clip in (2.5, 3.0):                                                   # cv2.createCLAHE(clip, (8, 8)).apply: bit-exact
        assert np.array_equal(fire.clahe_u8(u8, clip).cpu().numpy(), kat[f"clahe{clip}_" + tag]), clip
    for lo in (30, 50):                                                       # cv2.Canny(., lo, 150): bit-exact
        got = fire.canny_u8(u8, lo, 150).cpu().numpy()
        assert np.array_equal(got, kat[f"canny{lo}_" + tag]) and got.any(), lo
    g = torch.from_numpy(ref_fire.make_fire_frame(h, w, seed=h)[0]).to(cuda_device)
    assert np.array_equal(fire.histogram100(g).cpu().numpy(), kat["hist_" + tag])          # np.histogram: exact counts
    dx, dy = fire.sobel3(g)                                                   # cv2.Sobel float: one rounding
    np.testing.assert_allclose(dx.cpu().numpy(), kat["sobelx_" + tag], rtol=0, atol=1e-6)
    np.testing.assert_allclose(dy.cpu().numpy(), kat["sobely_" + tag], rtol=0, atol=1e-6)
    d = torch.from_numpy(kat["depth_" + tag]).to(cuda_device)
    np.testing.assert_allclose(fire.bilateral_filter(d, 5, 50, 50).cpu().numpy(), kat["bil5_" + tag], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("h,w", SIZES)
def test_reference_functions_reproduced(cuda_device, kat, h, w):
    """The three drop-in functions against the live reference's outputs on the same input, same NumPy seed."""
    from thermal3d_vision_b200 import fire
    tag = f"{h}x{w}"
    f = ref_fire.make_fire_frame(h, w, seed=h)
    np.random.seed(11)
    got = fire.preprocess_fire_scene_thermal(torch.from_numpy(f))
    assert not got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (3, h, w)
    np.testing.assert_allclose(got.numpy(), kat["pre_" + tag], rtol=0, atol=1e-6)
    assert (got.numpy() != kat["pre_" + tag]).mean() < 1e-3                    # bit-identical up to rare last-bit cases
    np.random.seed(11)
    on_dev = fire.preprocess_fire_scene_thermal(torch.from_numpy(f).to(cuda_device))
    assert on_dev.is_cuda and torch.equal(on_dev.cpu(), got)
    np.random.seed(12)
    adv = fire.advanced_fire_scene_processing(torch.from_numpy(f))
    np.testing.assert_allclose(adv.numpy(), kat["adv_" + tag], rtol=0, atol=2e-6)     # ends in the 9/75/75 bilateral filter
    ref = fire.depth_refinement_with_outlier_removal(kat["depth_" + tag], f, guided_filter=False)
    assert isinstance(ref, np.ndarray)
    np.testing.assert_allclose(ref, kat["refined_" + tag], rtol=2e-6, atol=2e-6)
    with pytest.raises(NotImplementedError):
        fire.depth_refinement_with_outlier_removal(kat["depth_" + tag], f)       # guided filter: no oracle here


@pytest.mark.parametrize("h,w", [(64, 64), (50, 83), (224, 224), (384, 512)])
def test_operators_match_restatements_on_other_shapes(cuda_device, h, w):
    """Shapes not in the golden file (tile grids that need the reflect-101 extension, the DUSt3R sizes) against the
    NumPy restatements of the library algorithms (pinned to the live libraries in tests/test_oracle_pin.py)."""
    from thermal3d_vision_b200 import fire
    rng = np.random.default_rng(h * w)
    f = ref_fire.make_fire_frame(h, w, seed=3)[0]
    u8 = np.clip(f * 255 + rng.integers(-6, 6, f.shape), 0, 255).astype(np.uint8)
    t = torch.from_numpy(u8).to(cuda_device)
    assert np.array_equal(fire.clahe_u8(t, 3.0).cpu().numpy(), ref_fire.clahe_u8(u8, 3.0))
    assert np.array_equal(fire.clahe_u8(t, 0.0).cpu().numpy(), ref_fire.clahe_u8(u8, 0.0))          # no clipping
    assert np.array_equal(fire.canny_u8(t, 50, 150).cpu().numpy(), ref_fire.canny_u8(u8, 50, 150))
    assert np.array_equal(fire.canny_u8(t, 150, 10).cpu().numpy(), ref_fire.canny_u8(u8, 10, 150))  # swapped thresholds
    x = rng.random((h, w)).astype(np.float32)
    x.flat[:5] = [0.0, 1.0, 0.01, 0.99, 0.57]                                                       # bin edges
    assert np.array_equal(fire.histogram100(torch.from_numpy(x).to(cuda_device)).cpu().numpy(), ref_fire.histogram100(x))
    d = (2 + rng.standard_normal((h, w))).astype(np.float32)
    d[rng.random((h, w)) < 0.02] -= 25
    from thermal3d_vision_b200 import _lib
    out = torch.empty(h, w, device=cuda_device); stats = torch.empty(2, device=cuda_device)
    mask = torch.empty(h, w, dtype=torch.uint8, device=cuda_device)
    dd = torch.from_numpy(d).to(cuda_device)
    _lib.check(_lib.lib().t3d_depth_outlier_median(_lib.ptr(dd), _lib.ptr(out), h, w, _lib.ptr(stats), _lib.ptr(mask),
                                                   _lib.current_stream_ptr()), "t3d_depth_outlier_median")
    ro, rm, mean, std = ref_fire.outlier_median(d)
    assert np.array_equal(mask.cpu().numpy().astype(bool), rm) and rm.any()
    np.testing.assert_allclose(stats.cpu().numpy(), [mean, std], rtol=1e-6)
    np.testing.assert_allclose(out.cpu().numpy(), ro, rtol=1e-6, atol=1e-6)
    assert (out.cpu().numpy()[~rm] == d[~rm]).all()
