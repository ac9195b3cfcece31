"""GPU parity: unsharp_mask (core/proc/unsharp_mask.cc:72-118; c_image_stacking_pipeline.cc:1302-1306) against
oracle/unsharp.py (cv2.sepFilter2D + cv2.addWeighted).

Tolerance: cv2's row / column filters use FMA in their vector body and separate multiply + add in the scalar tail that
handles the last cols % 8 columns, so the low-pass image is reproduced bit-exactly on the vector columns and to 1 ulp on
the tail columns; addWeighted (fp64 fma, one rounding to fp32) is reproduced exactly up to that input difference."""
import numpy as np
import pytest

from oracle import unsharp as ou

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("shape", [(64, 96), (97, 131), (270, 480), (120, 161, 3), (33, 40, 2), (1080, 1920)])
@pytest.mark.parametrize("sigma,alpha", [(1.0, 0.8), (1.0, 0.9), (0.5, 0.37), (2.0, 0.6), (1.5, 0.8)])
def test_unsharp_mask_matches_oracle(gpu, shape, sigma, alpha):
    from serstacker_b200 import api
    rng = np.random.default_rng(shape[1])
    src = rng.random(shape, dtype=f32)
    want = ou.unsharp_mask(src, sigma, alpha)
    got = api.unsharp_mask(src, sigma, alpha)
    beta = alpha / (1.0 - alpha)
    tol = (2.0 + beta) * 1.2e-7 * max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= tol
    body = shape[1] - shape[1] % 8                      # columns cv2 filters in its vector loops
    frac_exact = np.mean(got[:, :body] == want[:, :body])
    assert frac_exact >= 0.999, frac_exact


def test_unsharp_mask_clamp_copy_and_rejects(gpu):
    from serstacker_b200 import api
    rng = np.random.default_rng(3)
    src = rng.random((50, 64), dtype=f32)
    want = ou.unsharp_mask(src, 1.0, 0.8, 0.0, 1.0)
    got = api.unsharp_mask(src, 1.0, 0.8, 0.0, 1.0)
    assert got.min() >= 0.0 and got.max() <= 1.0 and (got == 0).any() and (got == 1).any()
    assert np.abs(got - want).max() <= 1e-6
    assert np.array_equal(api.unsharp_mask(src, 0.0, 0.8), src)       # sigma <= 0 or alpha <= 0: copy
    assert np.array_equal(api.unsharp_mask(src, 1.0, 0.0), src)
    with pytest.raises(Exception):
        api.unsharp_mask(src, 1.0, 1.0)                                  # alpha must stay below 1


@pytest.mark.parametrize("shape", [(64, 96), (97, 131), (270, 480), (1080, 1920)])
@pytest.mark.parametrize("sigma,alpha", [(2.5, 0.5), (3.0, 0.8), (4.0, 0.6), (6.0, 0.5), (10.0, 0.5)])
def test_unsharp_mask_pyramid_branch_matches_oracle(gpu, shape, sigma, alpha):
    """create_lpass_image's approximation for sigma > 2 (unsharp_mask.cc:50-68): pyrDown chain with BORDER_REFLECT, residual
    Gaussian, pyrUp chain through the size history."""
    from serstacker_b200 import api
    level, _ = ou.lpass_pyramid_level(shape[0], shape[1], sigma)
    assert level >= 1
    rng = np.random.default_rng(shape[1] + int(sigma * 10))
    src = rng.random(shape, dtype=f32)
    want = ou.unsharp_mask(src, sigma, alpha)
    got = api.unsharp_mask(src, sigma, alpha)
    beta = alpha / (1.0 - alpha)
    tol = (2.0 + beta) * 2.4e-7 * max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= tol


def test_unsharp_mask_pyramid_branch_rejects_colour(gpu):
    from serstacker_b200 import api
    with pytest.raises(Exception):
        api.unsharp_mask(np.zeros((64, 64, 3), f32), 3.0, 0.5)
