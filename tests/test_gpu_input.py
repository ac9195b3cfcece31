"""GPU parity: the input side of the loop (dark / flat calibration, average_bayer_planes, colour
matrix, linear_interpolation_inpaint) and the master-frame steps (rank 2: select_master_frame's metric, create_reference_frame)."""
import numpy as np
import cv2
import pytest

from oracle import debayer as od
from oracle import inpaint as oi
from oracle import pipeline as opl
from oracle import transforms as otf
from oracle import ecc as oecc
from serstacker_b200 import synth
from helpers import rel_l2

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("cn", [1, 3])
@pytest.mark.parametrize("hole_kind", ["random", "stripes", "border", "none", "all"])
def test_linear_interpolation_inpaint_matches_oracle(gpu, cn, hole_kind):
    from serstacker_b200 import api
    rng = np.random.default_rng(3)
    h, w = 97, 131
    img = rng.random((h, w) if cn == 1 else (h, w, cn)).astype(f32)
    mask = np.full((h, w), 255, np.uint8)
    if hole_kind == "random":
        mask[rng.random((h, w)) < 0.35] = 0
    elif hole_kind == "stripes":
        mask[10:14] = 0
        mask[:, 40:47] = 0
        mask[60:90, 100:] = 0
    elif hole_kind == "border":          # what a registered stack looks like: an invalid frame around the image
        mask[:6] = 0
        mask[-9:] = 0
        mask[:, :11] = 0
        mask[:, -4:] = 0
    elif hole_kind == "all":
        mask[:] = 0
    want = oi.linear_interpolation_inpaint(img, mask)
    got = api.linear_interpolation_inpaint(img, mask)
    assert got.shape == want.shape
    # the reference is built with -ffast-math (its own expressions may contract to FMAs): 1-ulp agreement
    assert np.abs(got - want).max() <= 2e-7 * max(1.0, float(np.abs(want).max())), np.abs(got - want).max()
    assert np.array_equal(got[mask > 0], img[mask > 0])
    assert np.array_equal(api.linear_interpolation_inpaint(img, None), img)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_average_bayer_planes_matches_oracle(gpu, dtype):
    from serstacker_b200 import api
    rng = np.random.default_rng(4)
    raw = rng.random((64, 96))
    raw = raw.astype(f32) if dtype == np.float32 else np.rint(raw * np.iinfo(dtype).max).astype(dtype)
    assert np.array_equal(api.average_bayer_planes(raw), od.average_bayer_planes(raw))
    with pytest.raises(Exception):
        api.average_bayer_planes(raw[:63])


@pytest.mark.parametrize("dtype,bpp", [(np.uint16, 16), (np.uint16, 12), (np.uint8, 8), (np.float32, 0)])
@pytest.mark.parametrize("cn", [1, 3])
def test_input_calibrate_matches_opencv(gpu, dtype, bpp, cn):
    """read_input_frame: convertTo(CV_32F, 1 / (1 << bpp)), cv::subtract(dark), cv::divide(flat) (c_image_stacking_pipeline_base.cc:143-184)."""
    from serstacker_b200 import api
    rng = np.random.default_rng(5)
    shape = (48, 80) if cn == 1 else (48, 80, cn)
    frame = rng.random(shape)
    frame = frame.astype(f32) if dtype == np.float32 else np.rint(frame * ((1 << bpp) - 1)).astype(dtype)
    dark = (rng.random(shape) * 0.05).astype(f32)
    flat = (0.5 + rng.random(shape)).astype(f32)
    flat.reshape(-1)[::97] = 0                       # cv::divide yields 0 where the divisor is 0
    ff = frame if dtype == np.float32 else cv2.convertScaleAbs(frame, alpha=1) if False else frame.astype(np.float64)
    want = frame.copy() if dtype == np.float32 else (frame.astype(np.float64) * (1.0 / (1 << bpp))).astype(f32)
    want = cv2.subtract(want, dark)
    want = cv2.divide(want, flat)
    got = api.input_calibrate(frame, bpp, dark, flat)
    assert np.abs(got - want).max() <= 1e-7 * max(1.0, float(np.abs(want).max()))
    only_dark = api.input_calibrate(frame, bpp, dark, None)
    base = frame.copy() if dtype == np.float32 else (frame.astype(np.float64) * (1.0 / (1 << bpp))).astype(f32)
    assert np.array_equal(only_dark, cv2.subtract(base, dark))


def test_color_transform_matches_opencv(gpu):
    from serstacker_b200 import api
    rng = np.random.default_rng(6)
    img = rng.random((40, 56, 3)).astype(f32)
    m = (np.eye(3) + 0.2 * rng.standard_normal((3, 3))).astype(f32)
    got = api.color_transform(img, m)
    want = cv2.transform(img, m)
    assert np.abs(got - want).max() <= 2e-7 * float(np.abs(want).max())


def test_select_master_frame_picks_the_sharpest(gpu):
    """master_frame_best_of_100_in_middle: the local-variance metric of every scanned frame, first maximum wins."""
    from serstacker_b200 import api
    frames, _, _ = synth.make_planet_sequence(320, 240, 7, seed=12, sigma_t=2.0, blur_range=(0.8, 3.0), dtype="f32")
    best_o, m_o = opl.select_master_frame(frames)
    best_g, m_g = api.select_master_frame(frames)
    assert best_g == best_o
    assert np.allclose(m_g, m_o, rtol=1e-5)
    raw, _, bpp = synth.make_bayer_sequence(128, 96, 4, seed=13)
    best_ob, m_ob = opl.select_master_frame(raw, bayer=True)
    best_gb, m_gb = api.select_master_frame(raw, colorid=8)
    # raw 16-bit frames: cv::pyrDown rounds the integer image at every level, the device chain runs in float
    assert best_gb == best_ob and np.allclose(m_gb, m_ob, rtol=2e-3)


def test_create_reference_frame_matches_oracle(gpu):
    """The master-frame pass: frames around the selected one stacked against it (REFLECT101 remap), compute(),
    linear_interpolation_inpaint, unsharp_mask(1, 0.8) (c_image_stacking_pipeline.cc:1112-1312)."""
    from serstacker_b200 import api
    frames, _, _ = synth.make_planet_sequence(320, 240, 9, seed=14, radius=70, sigma_t=3.0, dtype="f32")
    so = opl.StackingOptions()
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    ref_o, mask_o = opl.create_reference_frame(frames, 4, so, max_frames_to_stack=6)
    ro = api.registration_options(motion_type=0, ecc=dict(ecc_method=3, ecch_max_level=-1))
    go = api.stack_options(registration=ro, accumulation_method=0, max_batch=4)
    ref_g, mask_g = api.create_reference_frame(frames, 4, go, max_frames_to_stack=6)
    assert np.array_equal(mask_g, mask_o)
    assert rel_l2(ref_g, ref_o) <= 1e-5
    assert api.master_frame_range(100, 50, 30) == opl.master_frame_range(100, 50, 30) == (35, 65)
    assert api.master_frame_range(100, 95, 30) == opl.master_frame_range(100, 95, 30) == (70, 100)
    assert api.master_frame_range(20, 3, 3000) == (0, 20)
