"""GPU parity: the persistent ECC kernel (K2/K3/K4) against the oracle solvers, per method and transform."""
import numpy as np
import cv2
import pytest

from oracle import ecc as oecc
from oracle import transforms as otf
from oracle import registration as oreg
from serstacker_b200 import synth
from helpers import map_diff_px, dot_noise, strict_case

pytestmark = pytest.mark.gpu

METHODS = [oecc.ECC_ALIGN_FORWARD_ADDITIVE, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL, oecc.ECC_ALIGN_LM,
           oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM]


def _seq(w, h, n, seed, rot=0.0, scale=0.0, sigma_t=2.0):
    frames, mats, _ = synth.make_planet_sequence(w, h, n, seed, sigma_t=sigma_t, sigma_rot_deg=rot, sigma_scale=scale, dtype="f32")
    return frames, mats


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("motion", [0, 3, 1, 2, 4])
@pytest.mark.parametrize("maxlevel", [0, -1])
def test_ecch_align_matches_oracle(gpu, method, motion, maxlevel):
    """c_ecch::align on a half-resolution-like image: same trajectory => same parameters (<= 1e-3 px)."""
    from serstacker_b200 import api
    frames, _ = _seq(320, 240, 4, seed=11 + motion, rot=0.2 if motion else 0.0, scale=0.002 if motion in (2, 3, 4) else 0.0)
    kw = dict(maxlevel=maxlevel, minimum_image_size=16, epsx=0.05, max_iterations=30, update_step_scale=1.0)
    ot = otf.create_image_transform(motion)
    o = oecc.EccH(ot, method=method, **kw)
    o.set_reference_image(frames[0], None)
    gt = api.create_image_transform(motion)
    g = api.c_ecch(gt, method=method, **kw)
    g.set_reference_image(frames[0])
    strict = strict_case(motion, method)
    for f in frames[1:]:
        ot.reset()
        gt.set_parameters(ot.parameters())
        o.align(f, None)
        g.align(f)
        p_o = ot.parameters().copy()
        d = map_diff_px(motion, gt.parameters(), p_o, (320, 240))
        if strict:
            assert g.num_iterations() == o.num_iterations, (g.num_iterations(), o.num_iterations, d)
            assert d <= 1e-3, d
        else:
            # envelope of the reference algorithm under its own summation noise (see helpers.py)
            ot.reset()
            with dot_noise():
                o.align(f, None)
            env = map_diff_px(motion, ot.parameters(), p_o, (320, 240))
            assert d <= max(1e-3, (10 if motion == 1 else 4) * env), (d, env)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("motion,tfirst", [(0, True), (3, True), (3, False), (4, True), (1, True), (2, False)])
def test_register_frame_matches_oracle(gpu, method, motion, tfirst):
    """c_frame_registration::register_frame incl. scaleImage, translation-first, rho gate, scale back."""
    from serstacker_b200 import api
    frames, _ = _seq(400, 300, 5, seed=31 + motion, rot=0.15 if motion else 0.0, scale=0.002 if motion in (2, 3, 4) else 0.0, sigma_t=3.0)
    oo = oreg.ImageRegistrationOptions(motion_type=motion)
    oo.ecc.ecc_method = method
    oo.ecc.ecch_max_level = -1
    oo.ecc.ecch_estimate_translation_first = tfirst
    oo.ecc.update_step_scale = 1.0 if method in (oecc.ECC_ALIGN_LM,) else 1.5
    o = oreg.FrameRegistration(oo)
    o.setup_reference_frame(frames[0])
    go = api.registration_options(motion_type=motion, ecc=dict(ecc_method=method, ecch_max_level=-1,
                                                               ecch_estimate_translation_first=int(tfirst),
                                                               update_step_scale=oo.ecc.update_step_scale))
    g = api.c_frame_registration(go)
    g.setup_reference_frame(frames[0])
    strict = strict_case(motion, method)
    for f in frames:
        # a diverging trial that maps every pixel outside the image divides by CMA == 0: inf / NaN in the reference's double
        # arithmetic (ecc2.cc:1915, 1923), in the oracle (_ddiv) and on the device alike
        ok_o = o.register_frame(f)
        ok_g = g.register_frame(f)
        assert ok_o == ok_g
        assert (np.isnan(g.status.rho) and np.isnan(o.status.rho)) or abs(g.status.rho - o.status.rho) <= 1e-4
        if ok_o:
            p_o = o.image_transform.parameters().copy()
            d = map_diff_px(motion, g.image_transform_parameters(), p_o, (400, 300))
            if strict:
                assert d <= 1e-3, (d, g.status.num_iterations, o.status.num_iterations)
                assert g.status.num_iterations == o.status.num_iterations
            else:
                with dot_noise():
                    o.register_frame(f)
                env = map_diff_px(motion, o.image_transform.parameters(), p_o, (400, 300))
                # fixed-scale euclidean never meets its eps() test (c_image_transform.cc:509-523 adds max(w,h)*scale)
                # and runs all 50 over-relaxed iterations per level: allow a wider multiple of the envelope
                assert d <= max(1e-3, (10 if motion == 1 else 4) * env), (d, env)


@pytest.mark.parametrize("scale,size", [(0.25, (400, 300)), (0.25, (402, 302)), (0.75, (400, 300)), (0.4, (401, 299)), (1.0, (200, 150))])
def test_register_frame_with_ecc_scale(gpu, scale, size):
    """scaleImage's cv::resize(INTER_AREA) branch (c_frame_registration.cc:242-247): integer scale (ResizeAreaFast, with and
    without cells clipped by the image edge), fractional scale (ResizeArea), and scale 1 (no scaling)."""
    from serstacker_b200 import api
    w, h = size
    frames, _ = _seq(w, h, 4, seed=57, rot=0.0, scale=0.0, sigma_t=3.0)
    oo = oreg.ImageRegistrationOptions(motion_type=0)
    oo.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    oo.ecc.ecch_max_level = -1
    oo.ecc.scale = scale
    o = oreg.FrameRegistration(oo)
    o.setup_reference_frame(frames[0])
    g = api.c_frame_registration(api.registration_options(motion_type=0, ecc=dict(ecc_method=3, ecch_max_level=-1, scale=scale)))
    g.setup_reference_frame(frames[0])
    for f in frames:
        ok_o = o.register_frame(f)
        ok_g = g.register_frame(f)
        assert ok_o == ok_g
        assert abs(g.status.rho - o.status.rho) <= 1e-4
        if ok_o:
            d = map_diff_px(0, g.image_transform_parameters(), o.image_transform.parameters().copy(), (w, h))
            assert d <= 1e-3, (d, g.status.num_iterations, o.status.num_iterations)
            # the scaled ECC images are bit-identical to cv::resize's, so the translation solver takes the same steps
            assert g.status.num_iterations == o.status.num_iterations
            assert np.array_equal(g.image_transform_parameters(), o.image_transform.parameters().ravel()[:2])


def test_low_correlation_frame_is_dropped(gpu):
    from serstacker_b200 import api
    frames, _ = _seq(320, 240, 2, seed=5)
    rng = np.random.default_rng(0)
    junk = rng.random((240, 320)).astype(np.float32)
    go = api.registration_options(motion_type=0, ecc=dict(ecc_method=oecc.ECC_ALIGN_FORWARD_ADDITIVE))
    g = api.c_frame_registration(go)
    g.setup_reference_frame(frames[0])
    assert g.register_frame(frames[1]) is True
    assert g.register_frame(junk) is False
    oo = oreg.ImageRegistrationOptions(motion_type=0)
    oo.ecc.ecc_method = oecc.ECC_ALIGN_FORWARD_ADDITIVE
    o = oreg.FrameRegistration(oo)
    o.setup_reference_frame(frames[0])
    assert o.register_frame(junk) is False


def _disk_mask(w, h, rx, ry, cx=None, cy=None):
    """Elliptical reference mask (planetary disk ROI), CV_8UC1 0/255."""
    cx = (w - 1) / 2 if cx is None else cx
    cy = (h - 1) / 2 if cy is None else cy
    yy, xx = np.mgrid[0:h, 0:w]
    return ((((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0) * 255).astype(np.uint8)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("motion", [0, 3])
def test_ecch_reference_mask_matches_oracle(gpu, method, motion):
    """c_ecch::set_reference_image(image, mask): mask pyramid by INTER_NEAREST, per-solver erosion, masked reference
    gradients, RMA = countNonZero (ecc2.cc:972-1057, 1211-1221, 1425-1431, 1855-1860)."""
    from serstacker_b200 import api
    frames, _ = _seq(320, 240, 3, seed=51 + motion, rot=0.2 if motion else 0.0, scale=0.002 if motion else 0.0)
    mask = _disk_mask(320, 240, 120, 90, cx=165, cy=118)
    kw = dict(maxlevel=-1, minimum_image_size=16, epsx=0.05, max_iterations=30, update_step_scale=1.0)
    ot = otf.create_image_transform(motion)
    o = oecc.EccH(ot, method=method, **kw)
    o.set_reference_image(frames[0], mask)
    gt = api.create_image_transform(motion)
    g = api.c_ecch(gt, method=method, **kw)
    g.set_reference_image(frames[0], mask)
    for f in frames[1:]:
        ot.reset()
        gt.set_parameters(ot.parameters())
        o.align(f, None)
        g.align(f)
        p_o = ot.parameters().copy()
        d = map_diff_px(motion, gt.parameters(), p_o, (320, 240))
        if strict_case(motion, method):
            assert g.num_iterations() == o.num_iterations, (g.num_iterations(), o.num_iterations, d)
            assert d <= 1e-3, d
        else:
            ot.reset()
            with dot_noise():
                o.align(f, None)
            env = map_diff_px(motion, ot.parameters(), p_o, (320, 240))
            assert d <= max(1e-3, 4 * env), (d, env)
    # the mask must matter: the unmasked alignment of the last frame follows a different trajectory
    o2 = oecc.EccH(otf.create_image_transform(motion), method=method, **kw)
    o2.set_reference_image(frames[0], None)
    o2.align(frames[-1], None)
    assert not np.array_equal(o2.transform.parameters(), p_o)


@pytest.mark.parametrize("method", [oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM, oecc.ECC_ALIGN_LM])
def test_register_frame_with_reference_mask(gpu, method):
    """setup_reference_frame(image, mask): scaleImage's mask branch (pyrDown of the 8-bit mask, >= 250) feeds c_ecch
    (c_frame_registration.cc:230-250, 565-719); the correlation gate uses the same mask (ecc2.cc:65-137)."""
    from serstacker_b200 import api
    frames, _ = _seq(401, 300, 4, seed=77, rot=0.15, scale=0.002, sigma_t=3.0)
    mask = _disk_mask(401, 300, 150, 110)
    oo = oreg.ImageRegistrationOptions(motion_type=3)
    oo.ecc.ecc_method = method
    oo.ecc.ecch_max_level = -1
    oo.ecc.update_step_scale = 1.0 if method == oecc.ECC_ALIGN_LM else 1.5
    o = oreg.FrameRegistration(oo)
    o.setup_reference_frame(frames[0], mask)
    g = api.c_frame_registration(api.registration_options(motion_type=3, ecc=dict(
        ecc_method=method, ecch_max_level=-1, update_step_scale=oo.ecc.update_step_scale)))
    g.setup_reference_frame(frames[0], mask)
    for f in frames:
        ok_o = o.register_frame(f)
        ok_g = g.register_frame(f)
        assert ok_o == ok_g
        assert abs(g.status.rho - o.status.rho) <= 1e-4
        if ok_o:
            d = map_diff_px(3, g.image_transform_parameters(), o.image_transform.parameters(), (401, 300))
            assert d <= 1e-3, (d, g.status.num_iterations, o.status.num_iterations)


@pytest.mark.parametrize("method", [oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM, oecc.ECC_ALIGN_FORWARD_ADDITIVE])
@pytest.mark.parametrize("nscale,size", [(2, (400, 300)), (3, (401, 299))])
def test_register_frame_with_ecc_normalize(gpu, method, nscale, size):
    """ecc.normalization_scale > 0: ecc_normalize (pyrDown BORDER_REPLICATE chain, pyrUp back, subtract; ecc2.cc:385-397)
    on the reference and on every current ECC image (c_frame_registration.cc:640-660, 790-810)."""
    from serstacker_b200 import api
    w, h = size
    frames, _ = _seq(w, h, 4, seed=91, rot=0.1, scale=0.001, sigma_t=2.5)
    mask = _disk_mask(w, h, w * 0.4, h * 0.4)
    oo = oreg.ImageRegistrationOptions(motion_type=3)
    oo.ecc.ecc_method = method
    oo.ecc.ecch_max_level = -1
    oo.ecc.normalization_scale = nscale
    oo.ecc.normalization_noise = 0.01
    oo.ecc.update_step_scale = 1.0
    o = oreg.FrameRegistration(oo)
    o.setup_reference_frame(frames[0], mask)
    g = api.c_frame_registration(api.registration_options(motion_type=3, ecc=dict(
        ecc_method=method, ecch_max_level=-1, normalization_scale=nscale, normalization_noise=0.01, update_step_scale=1.0)))
    g.setup_reference_frame(frames[0], mask)
    for f in frames:
        ok_o = o.register_frame(f)
        ok_g = g.register_frame(f)
        assert ok_o == ok_g
        assert abs(g.status.rho - o.status.rho) <= 1e-4
        if ok_o:
            p_o = o.image_transform.parameters().copy()
            d = map_diff_px(3, g.image_transform_parameters(), p_o, size)
            if method == oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM:
                assert d <= 1e-3, (d, g.status.num_iterations, o.status.num_iterations)
            else:
                with dot_noise():
                    o.register_frame(f)
                env = map_diff_px(3, o.image_transform.parameters(), p_o, size)
                assert d <= max(1e-3, 4 * env), (d, env)


def _current_mask(shape, kind):
    h, w = shape
    m = np.full((h, w), 255, np.uint8)
    if kind == "blob":                   # bad-pixel islands and a saturated region, as read_input_frame's masks have them
        m[40:70, 100:160] = 0
        m[150:152, 30:200] = 0
        m[::37, ::41] = 0
    elif kind == "roi":
        m[:, :25] = 0
        m[-18:] = 0
    return m


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("motion", [0, 3])
@pytest.mark.parametrize("kind", ["blob", "roi"])
def test_ecch_align_with_current_mask_matches_oracle(gpu, method, motion, kind):
    """c_ecch::align(current_image, current_mask): the mask pyramid (resize NEAREST), the forward-additive erosion, and the
    per-solver remap of the mask (LINEAR >= 255 / >= 250, NEAREST of the inverted mask for IC-LM; ecc2.cc:1307-1316,
    205-216, 1877-1884)."""
    from serstacker_b200 import api
    frames, _ = _seq(320, 240, 4, seed=71 + motion, rot=0.2 if motion else 0.0, scale=0.002 if motion else 0.0)
    kw = dict(maxlevel=-1, minimum_image_size=16, epsx=0.05, max_iterations=30, update_step_scale=1.0)
    ot = otf.create_image_transform(motion)
    o = oecc.EccH(ot, method=method, **kw)
    o.set_reference_image(frames[0], None)
    gt = api.create_image_transform(motion)
    g = api.c_ecch(gt, method=method, **kw)
    g.set_reference_image(frames[0])
    mask = _current_mask(frames[0].shape, kind)
    strict = strict_case(motion, method)
    for f in frames[1:]:
        ot.reset()
        gt.set_parameters(ot.parameters())
        o.align(f, mask)
        g.align(f, mask)
        p_o = ot.parameters().copy()
        d = map_diff_px(motion, gt.parameters(), p_o, (320, 240))
        # the mask must matter: without it the oracle lands elsewhere (or takes another number of iterations)
        if strict:
            assert g.num_iterations() == o.num_iterations, (g.num_iterations(), o.num_iterations, d)
            assert d <= 1e-3, d
        else:
            ot.reset()
            with dot_noise():
                o.align(f, mask)
            env = map_diff_px(motion, ot.parameters(), p_o, (320, 240))
            assert d <= max(1e-3, 4 * env), (d, env)
    # an unmasked align after a masked one is not affected by the previous mask
    ot.reset()
    gt.set_parameters(ot.parameters())
    o.align(frames[1], None)
    g.align(frames[1])
    if strict:
        assert map_diff_px(motion, gt.parameters(), ot.parameters(), (320, 240)) <= 1e-3


@pytest.mark.parametrize("method", [oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM, oecc.ECC_ALIGN_FORWARD_ADDITIVE, oecc.ECC_ALIGN_LM])
def test_register_frame_with_current_mask(gpu, method):
    """c_frame_registration::register_frame(src, srcmask): scaleImage of the mask (pyrDown >= 250), masked alignment, and the
    correlation gate under the remapped current mask (>= 254)."""
    from serstacker_b200 import api
    frames, _ = _seq(400, 300, 4, seed=83, rot=0.0, scale=0.0, sigma_t=3.0)
    oo = oreg.ImageRegistrationOptions(motion_type=0)
    oo.ecc.ecc_method = method
    oo.ecc.ecch_max_level = -1
    oo.ecc.update_step_scale = 1.0 if method == oecc.ECC_ALIGN_LM else 1.5
    o = oreg.FrameRegistration(oo)
    o.setup_reference_frame(frames[0])
    g = api.c_frame_registration(api.registration_options(motion_type=0, ecc=dict(ecc_method=method, ecch_max_level=-1,
                                                                                    update_step_scale=oo.ecc.update_step_scale)))
    g.setup_reference_frame(frames[0])
    mask = _current_mask(frames[0].shape, "blob")
    mask[200:260, 250:330] = 0
    for f in frames[1:]:
        ok_o = o.register_frame(f, mask)
        ok_g = g.register_frame(f, mask)
        assert ok_o == ok_g
        assert abs(g.status.rho - o.status.rho) <= 1e-4
        if ok_o:
            assert map_diff_px(0, g.image_transform_parameters(), o.image_transform.parameters(), (400, 300)) <= 1e-3
            assert g.status.num_iterations == o.status.num_iterations
