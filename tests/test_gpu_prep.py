"""GPU parity: K1 (ECC image preparation, pyramids) and K6 (W1 weight map) against the oracle / cv2."""
import numpy as np
import cv2
import pytest

from oracle import ecc as oecc
from oracle import weights as ow
from serstacker_b200 import synth

pytestmark = pytest.mark.gpu


def _frame(w, h, seed, dtype="f32"):
    frames, _, bpp = synth.make_planet_sequence(w, h, 1, seed, sigma_t=0, dtype=dtype)
    return frames[0], bpp


@pytest.mark.parametrize("size", [(320, 240), (333, 251), (960, 540)])
@pytest.mark.parametrize("sigma", [1.0, 0.0, 1.7])
def test_reference_pyramid_matches_oracle(gpu, size, sigma):
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 3)
    o = oecc.EccH(oecc_transform(), maxlevel=-1, minimum_image_size=16, reference_smooth_sigma=sigma)
    o.set_reference_image(img, None)
    g = api.c_ecch(None, maxlevel=-1, minimum_image_size=16, reference_smooth_sigma=sigma)
    g.set_reference_image(img)
    assert g.num_levels() == len(o.pyramid)
    for l, e in enumerate(o.pyramid):
        want = e.reference_image
        assert g.level_size(l) == (want.shape[1], want.shape[0])
        got = g.reference_image(l)
        assert np.abs(got - want).max() <= 2e-6, (l, np.abs(got - want).max())


def oecc_transform():
    from oracle import transforms as otf
    return otf.TranslationTransform()


def test_single_level_default(gpu):
    """ecch_max_level = 0 yields a single-level pyramid (ecc2.cc:1015)."""
    from serstacker_b200 import api
    img, _ = _frame(200, 160, 4)
    g = api.c_ecch(None, maxlevel=0)
    g.set_reference_image(img)
    assert g.num_levels() == 1


@pytest.mark.parametrize("size", [(320, 240), (301, 203), (1920, 1080)])
@pytest.mark.parametrize("kradius,dscale", [(1, 1), (2, 1), (1, 0), (1, 2)])
def test_local_variance_map_matches_oracle(gpu, size, kradius, dscale):
    from serstacker_b200 import api
    if size[0] > 1000 and (kradius, dscale) != (1, 1):
        pytest.skip("full size only for the default options")
    img, _ = _frame(size[0], size[1], 7)
    Qo, Mo = ow.compute_local_variance_map(img, dscale=dscale, kradius=kradius, uscale=0)
    Qg, Mg = api.compute_local_variance_map(img, dscale=dscale, kradius=kradius, uscale=0)
    assert abs(Qg - Qo) <= 2e-5 * abs(Qo)
    scale = np.abs(Mo).max()
    assert np.abs(Mg - Mo).max() <= 5e-6 * scale, np.abs(Mg - Mo).max() / scale


@pytest.mark.parametrize("size", [(320, 240), (301, 203), (652, 490)])
@pytest.mark.parametrize("dscale,uscale", [(1, 1), (1, 2), (0, 1), (2, 3)])
def test_local_variance_map_uscale_matches_oracle(gpu, size, dscale, uscale):
    """uscale > 0: cv::resize(INTER_AREA) of the map to dscaleSize(size, uscale) before the offset and the up-sampling
    (c_local_variance_sharpness_measure.cc:231-234); integer and fractional area ratios."""
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 8)
    Qo, Mo = ow.compute_local_variance_map(img, dscale=dscale, kradius=1, uscale=uscale)
    Qg, Mg = api.compute_local_variance_map(img, dscale=dscale, kradius=1, uscale=uscale)
    assert abs(Qg - Qo) <= 2e-5 * abs(Qo)
    scale = np.abs(Mo).max()
    assert np.abs(Mg - Mo).max() <= 5e-6 * scale, np.abs(Mg - Mo).max() / scale


def test_level0_smoothing_bit_exact(gpu):
    """The 7-tap Gaussian sepFilter2D (and the 5/3-tap derivative filters) follow OpenCV's filter-engine
    arithmetic exactly, so the reference-side Hessian sees the same gradients as the oracle."""
    from serstacker_b200 import api
    img, _ = _frame(320, 240, 9)
    o = oecc.EccH(oecc_transform(), maxlevel=0, reference_smooth_sigma=1.0)
    o.set_reference_image(img, None)
    g = api.c_ecch(None, maxlevel=0, reference_smooth_sigma=1.0)
    g.set_reference_image(img)
    assert np.array_equal(g.reference_image(0), o.pyramid[0].reference_image)


@pytest.mark.parametrize("size", [(320, 240), (336, 252), (960, 540)])
def test_pyramid_bit_exact(gpu, size):
    """Gaussian smoothing + the whole cv::pyrDown chain are bit-identical to cv2, level by level (level-0 widths
    that are a multiple of 4; OpenCV's scalar tail columns of sepFilter2D are matched to 2e-6 only)."""
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 5)
    o = oecc.EccH(oecc_transform(), maxlevel=-1, minimum_image_size=16)
    o.set_reference_image(img, None)
    g = api.c_ecch(None, maxlevel=-1, minimum_image_size=16)
    g.set_reference_image(img)
    for l, e in enumerate(o.pyramid):
        assert np.array_equal(g.reference_image(l), e.reference_image), l


@pytest.mark.parametrize("size", [(320, 240), (333, 251), (130, 97), (161, 82)])
@pytest.mark.parametrize("k,p,dscale,uscale", [(2.0, 2.0, 2, 6), (2.0, 1.0, 0, 0), (1.0, 3.0, 1, 3), (3.0, 2.0, 2, 2)])
def test_lpg_matches_oracle(gpu, size, k, p, dscale, uscale):
    """W2: lpg (lpg.cc:223-290) = pdownscale, 5x5 Laplacian/gradient energy, pdownscale, pow, pyrUp chain.
    (130, 97) and (161, 82): last tile of the fused dscale = 0 kernel one or two pixels wide / high (its stencil centres are
    clamped into the neighbouring tile)."""
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 5)
    want = ow.lpg(img, k, p, dscale, uscale)
    got = api.lpg(img, k, p, dscale, uscale)
    assert got.shape == want.shape
    scale = float(np.abs(want).max())
    assert np.abs(got - want).max() <= 2e-6 * scale, (np.abs(got - want).max(), scale)


@pytest.mark.parametrize("cn,dtype", [(3, np.float32), (3, np.uint16), (2, np.float32), (4, np.uint8)])
def test_lpg_colour_images_average_their_channels(gpu, cn, dtype):
    """lpg of a multi-channel image: reduce_color_channels(REDUCE_AVG) first (lpg.cc:246-248)."""
    from serstacker_b200 import api
    rng = np.random.default_rng(8)
    base, _ = _frame(200, 144, 9)
    img = np.stack([np.clip(base * (0.6 + 0.2 * c) + rng.normal(0, 0.01, base.shape), 0, 1) for c in range(cn)], axis=2).astype(np.float32)
    if dtype != np.float32:
        img = np.rint(img * np.iinfo(dtype).max).astype(dtype)
    want = ow.lpg(img, 6.0, 2.0, 0, 0)
    got = api.lpg(img, 6.0, 2.0, 0, 0)
    scale = float(np.abs(want).max())
    assert np.abs(got - want).max() <= 2e-6 * scale, (np.abs(got - want).max(), scale)


def test_lpg_rejects_fractional_power(gpu):
    from serstacker_b200 import api, capi
    img, _ = _frame(64, 48, 6)
    with pytest.raises(capi.SskError):
        api.lpg(img, 2.0, 1.5, 1, 2)


@pytest.mark.parametrize("size", [(320, 240), (333, 251)])
@pytest.mark.parametrize("sigma", [1.0, 1.7])
def test_gaussian_blur_matches_opencv(gpu, size, sigma):
    """W3: cv::GaussianBlur(weights, Size(), sigma, sigma, BORDER_REPLICATE) of c_jdr_pipeline.cc:1228."""
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 7)
    want = cv2.GaussianBlur(img, (0, 0), sigma, None, sigma, cv2.BORDER_REPLICATE)
    got = api.gaussian_blur(img, sigma, sigma)
    assert np.abs(got - want).max() <= 2e-7 * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("size", [(320, 240), (1920, 1080), (648, 486), (328, 250), (264, 34)])
def test_w1_upsample2x_wide_kernel_is_bit_identical(gpu, size):
    """The eight-outputs-per-thread 2x up-sampling (analytic fractions + neighbour shuffles away from the edges) against the
    table-driven four-wide kernel it replaces in the stacking loop (SSK_W1_UP_V4 selects the latter)."""
    import os
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 11)
    Qa, Ma = api.compute_local_variance_map(img, dscale=1, kradius=1, uscale=0)
    os.environ["SSK_W1_UP_V4"] = "1"
    try:
        Qb, Mb = api.compute_local_variance_map(img, dscale=1, kradius=1, uscale=0)
    finally:
        del os.environ["SSK_W1_UP_V4"]
    assert Qa == Qb and np.array_equal(Ma, Mb)


@pytest.mark.parametrize("size", [(480, 400), (301, 203), (1024, 1024), (130, 95)])
@pytest.mark.parametrize("k,p,dscale,uscale", [(2.0, 2.0, 2, 6), (1.0, 3.0, 1, 3), (2.0, 2.0, 0, 5), (2.0, 1.0, 1, 8)])
def test_lpg_pyramid_kernels_are_bit_identical(gpu, size, k, p, dscale, uscale):
    """lpg's pyramid through the four-outputs-per-thread pyrUp (k_pyrup_2x2) and the single cluster launch for the small levels
    (k_lpg_tail) against the per-level, one-thread-per-output launches (SSK_PYRUP_V1, SSK_LPG_NO_TAIL): same bits on even, odd
    and tiny levels, and within the oracle's tolerance."""
    import os
    from serstacker_b200 import api
    img, _ = _frame(size[0], size[1], 13)
    a = api.lpg(img, k, p, dscale, uscale)
    os.environ["SSK_PYRUP_V1"] = "1"
    os.environ["SSK_LPG_NO_TAIL"] = "1"
    try:
        b = api.lpg(img, k, p, dscale, uscale)
    finally:
        del os.environ["SSK_PYRUP_V1"]
        del os.environ["SSK_LPG_NO_TAIL"]
    assert np.array_equal(a, b)
    want = ow.lpg(img, k, p, dscale, uscale)
    assert np.abs(a - want).max() <= 2e-6 * np.abs(want).max()
