"""GPU parity: cv::remap equivalents (K5 building blocks) against cv2 / the oracle."""
import numpy as np
import cv2
import pytest

from oracle import transforms as otf
from oracle import registration as oreg

pytestmark = pytest.mark.gpu

BORDERS = [cv2.BORDER_REFLECT101, cv2.BORDER_REPLICATE, cv2.BORDER_CONSTANT, cv2.BORDER_REFLECT, cv2.BORDER_WRAP]
INTERPS = [cv2.INTER_LINEAR, cv2.INTER_CUBIC, cv2.INTER_NEAREST, cv2.INTER_LANCZOS4]


def _rand_transform(rng, motion):
    from serstacker_b200 import api
    t = api.create_image_transform(motion)
    o = otf.create_image_transform(motion)
    if motion == otf.IMAGE_MOTION_TRANSLATION:
        p = rng.normal(0, 4, 2)
    elif motion == otf.IMAGE_MOTION_EUCLIDEAN:
        p = [rng.normal(0, 4), rng.normal(0, 4), rng.normal(0, 0.02)]
    elif motion == otf.IMAGE_MOTION_SCALED_EUCLIDEAN:
        p = [rng.normal(0, 4), rng.normal(0, 4), rng.normal(0, 0.02), 1 + rng.normal(0, 0.01)]
    elif motion == otf.IMAGE_MOTION_AFFINE:
        p = [1 + rng.normal(0, 0.01), rng.normal(0, 0.01), rng.normal(0, 4), rng.normal(0, 0.01), 1 + rng.normal(0, 0.01), rng.normal(0, 4)]
    else:
        p = [1 + rng.normal(0, 0.01), rng.normal(0, 0.01), rng.normal(0, 4), rng.normal(0, 0.01), 1 + rng.normal(0, 0.01), rng.normal(0, 4),
             rng.normal(0, 1e-5), rng.normal(0, 1e-5)]
    p = np.asarray(p, np.float32)
    t.set_parameters(p)
    o.set_parameters(p)
    return t, o


@pytest.mark.parametrize("motion", [0, 1, 2, 3, 4])
def test_create_remap_bit_exact(gpu, motion):
    rng = np.random.default_rng(motion)
    for _ in range(3):
        t, o = _rand_transform(rng, motion)
        got = t.create_remap((97, 61))
        want = o.create_remap((97, 61))
        if motion in (0, 3):
            assert np.array_equal(got, want)
        else:   # sin/cos (euclidean) and the division (homography) may differ in the last ulp
            assert np.abs(got - want).max() <= 2e-5


@pytest.mark.parametrize("interp", INTERPS)
@pytest.mark.parametrize("border", BORDERS)
def test_remap_matches_cv2(gpu, interp, border):
    from serstacker_b200 import api
    rng = np.random.default_rng(10 * interp + border)
    src = rng.random((83, 131)).astype(np.float32)
    for motion in (0, 3, 4):
        t, o = _rand_transform(rng, motion)
        rmap = o.create_remap((131, 83))
        want = cv2.remap(src, rmap, None, interp, borderMode=border, borderValue=0.25)
        got, _ = api.remap(t, None, src, interpolation=interp, border_mode=border, border_value=(0.25, 0, 0, 0))
        d = np.abs(got - want)
        # a coordinate that differs in the last ulp may fall into the neighbouring 1/32-px bucket: allow a
        # handful of such pixels (none for the exactly-representable translation / affine maps)
        nbad = int((d > 2e-6).sum())
        assert nbad <= (0 if motion in (0, 3) else 8), (motion, nbad, d.max())
        got2, _ = api.remap(None, rmap, src, interpolation=interp, border_mode=border, border_value=(0.25, 0, 0, 0))
        assert np.abs(got2 - want).max() <= 2e-6


def test_remap_lanczos4_bit_exact_and_transparent(gpu):
    """ECC_INTER_LANCZOS4 (ecc2.h:38): bit-exact against cv::remap, also at the border and with BORDER_TRANSPARENT."""
    from serstacker_b200 import api
    rng = np.random.default_rng(44)
    src = rng.random((83, 131)).astype(np.float32)
    t, o = _rand_transform(rng, 3)
    rmap = o.create_remap((131, 83))
    for border in BORDERS:
        want = cv2.remap(src, rmap, None, cv2.INTER_LANCZOS4, borderMode=border, borderValue=0.25)
        got, _ = api.remap(t, None, src, interpolation=cv2.INTER_LANCZOS4, border_mode=border, border_value=(0.25, 0, 0, 0))
        assert np.array_equal(got, want), (border, np.abs(got - want).max())
    dst0 = rng.random((83, 131)).astype(np.float32)
    want = cv2.remap(src, rmap, None, cv2.INTER_LANCZOS4, dst=dst0.copy(), borderMode=cv2.BORDER_TRANSPARENT)
    got, _ = api.remap(None, rmap, src, interpolation=cv2.INTER_LANCZOS4, border_mode=cv2.BORDER_TRANSPARENT, dst=dst0.copy())
    assert np.array_equal(got, want)
    src3 = rng.random((50, 70, 3)).astype(np.float32)
    rmap3 = o.create_remap((70, 50))
    assert np.array_equal(api.remap(None, rmap3, src3, interpolation=cv2.INTER_LANCZOS4, border_mode=cv2.BORDER_REFLECT101)[0],
                          cv2.remap(src3, rmap3, None, cv2.INTER_LANCZOS4, borderMode=cv2.BORDER_REFLECT101))


def test_remap_inter_area_is_linear(gpu):
    """cv::remap replaces INTER_AREA by INTER_LINEAR (ECC_INTER_AREA, ecc2.h:37): same through the C ABI."""
    from serstacker_b200 import api
    rng = np.random.default_rng(77)
    src = rng.random((83, 131)).astype(np.float32)
    t, o = _rand_transform(rng, 3)
    rmap = o.create_remap((131, 83))
    want = cv2.remap(src, rmap, None, cv2.INTER_AREA, borderMode=cv2.BORDER_REFLECT101)
    got, _ = api.remap(t, None, src, interpolation=cv2.INTER_AREA, border_mode=cv2.BORDER_REFLECT101, border_value=(0, 0, 0, 0))
    assert np.array_equal(got, want)


def test_remap_transparent_linear(gpu):
    from serstacker_b200 import api
    rng = np.random.default_rng(5)
    src = rng.random((60, 80)).astype(np.float32)
    t, o = _rand_transform(rng, 3)
    rmap = o.create_remap((80, 60))
    rmap[10:20, 10:30] = -1.0           # hidden-side convention of the derotation remap
    dst0 = rng.random((60, 80)).astype(np.float32)
    want = cv2.remap(src, rmap, None, cv2.INTER_LINEAR, dst=dst0.copy(), borderMode=cv2.BORDER_TRANSPARENT)
    got, _ = api.remap(None, rmap, src, interpolation=cv2.INTER_LINEAR, border_mode=cv2.BORDER_TRANSPARENT, dst=dst0.copy())
    assert np.abs(got - want).max() <= 2e-6


def test_remap_multichannel(gpu):
    from serstacker_b200 import api
    rng = np.random.default_rng(6)
    src = rng.random((50, 70, 3)).astype(np.float32)
    t, o = _rand_transform(rng, 3)
    rmap = o.create_remap((70, 50))
    for interp in (cv2.INTER_LINEAR, cv2.INTER_CUBIC):
        want = cv2.remap(src, rmap, None, interp, borderMode=cv2.BORDER_REFLECT101)
        got, _ = api.remap(t, None, src, interpolation=interp, border_mode=cv2.BORDER_REFLECT101)
        assert np.abs(got - want).max() <= 2e-6


@pytest.mark.parametrize("interp", INTERPS)
def test_remap_mask_matches_base_remap(gpu, interp):
    """mask = erode5x5(remap(all-255, interp, CONSTANT 0) >= 255, border 255): c_frame_registration.cc:1321-1338."""
    from serstacker_b200 import api
    rng = np.random.default_rng(20 + interp)
    src = rng.random((90, 120)).astype(np.float32)
    reg = oreg.FrameRegistration(oreg.ImageRegistrationOptions())
    for motion in (0, 3, 3, 4):
        t, o = _rand_transform(rng, motion)
        rmap = o.create_remap((120, 90))
        _, want = reg.base_remap(rmap, src, None, interpolation=interp, border_mode=cv2.BORDER_REFLECT101)
        _, got = api.remap(t, None, src, want_mask=True, interpolation=interp)
        assert int((got != want).sum()) <= (0 if motion in (0, 3) else 4)
    # arbitrary user mask through the fixed-point path
    m = np.full((90, 120), 255, np.uint8)
    m[30:40, 50:70] = 0
    m[rng.random((90, 120)) < 0.01] = 0
    t, o = _rand_transform(rng, 3)
    rmap = o.create_remap((120, 90))
    _, want = reg.base_remap(rmap, src, m, interpolation=interp, border_mode=cv2.BORDER_REFLECT101)
    _, got = api.remap(t, None, src, want_mask=True, src_mask=m, interpolation=interp)
    assert np.array_equal(got, want)


def test_bilinear_remap_bit_exact(gpu):
    """Bilinear sampling reproduces cv::remap's scalar float path bit for bit (translation / affine maps)."""
    from serstacker_b200 import api
    rng = np.random.default_rng(77)
    src = rng.random((83, 131)).astype(np.float32)
    for motion in (0, 3):
        for border in (cv2.BORDER_REPLICATE, cv2.BORDER_CONSTANT, cv2.BORDER_REFLECT101):
            t, o = _rand_transform(rng, motion)
            want = cv2.remap(src, o.create_remap((131, 83)), None, cv2.INTER_LINEAR, borderMode=border, borderValue=0.0)
            got, _ = api.remap(t, None, src, interpolation=cv2.INTER_LINEAR, border_mode=border)
            assert np.array_equal(got, want)
