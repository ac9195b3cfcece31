"""GPU parity: frame up-scaling of c_image_stacking_pipeline (upscale_image / upscale_remap / upscale_optflow,
c_image_stacking_pipeline.cc:1869-2002) and the stacking loop with frame_upscale_after_align (:1633-1660).

cv2.pyrUp is reproduced bit for bit; cv::resize(INTER_LINEAR) on CV_32F as a non-IPP OpenCV build computes it (the oracle
switches IPP off for these calls: the reference links the distribution's library) to <= 1 ulp (99.9 % of the samples bit-exact)."""
import numpy as np
import cv2
import pytest

from oracle import pipeline as opl
from oracle import ecc as oecc
from oracle import transforms as otf

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("option", [1, 2, 3])
@pytest.mark.parametrize("shape", [(37, 53), (64, 96, 3), (135, 241)])
def test_upscale_image_matches_oracle(gpu, option, shape):
    from serstacker_b200 import api
    rng = np.random.default_rng(option * 100 + shape[1])
    src = rng.standard_normal(shape).astype(f32)
    mask = np.full(shape[:2], 255, np.uint8)
    mask[5:20, 10:30] = 0
    mask[::11, ::7] = 0
    mask[:, -1] = 0
    want, wmask = opl.upscale_image(option, src, mask)
    got, gmask = api.upscale_image(option, src, mask)
    assert got.shape == want.shape and gmask.shape == wmask.shape
    if option == 1:
        assert np.array_equal(got, want)
    else:
        assert np.abs(got - want).max() <= 2.4e-7 * max(1.0, float(np.abs(want).max())) and (got == want).mean() > 0.99
    assert np.array_equal(gmask, wmask)


@pytest.mark.parametrize("option", [0, 1, 2, 3])
def test_upscale_remap_and_optflow_match_oracle(gpu, option):
    from serstacker_b200 import api
    h, w = 90, 140
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    rng = np.random.default_rng(option)
    rmap = np.stack([xx * f32(1.002) + yy * f32(0.004) + f32(1.7), yy * f32(0.998) - xx * f32(0.003) - f32(2.2)], -1).astype(f32)
    rmap += cv2.GaussianBlur(rng.standard_normal((h, w, 2)).astype(f32), (0, 0), 6.0)
    want, got = opl.upscale_remap(option, rmap), api.upscale_remap(option, rmap)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= (0 if option in (0, 1) else 2e-5)          # one ulp of a coordinate ~ 150
    flow = (rmap - np.stack([xx, yy], -1)).astype(f32)
    want, got = opl.upscale_optflow(option, flow), api.upscale_optflow(option, flow)
    assert np.abs(got - want).max() <= (0 if option in (0, 1) else 1e-6)


def _sequence(n=5, size=(320, 240)):
    from serstacker_b200 import synth
    frames, _, _ = synth.make_planet_sequence(size[0], size[1], n, seed=4, radius=min(size) * 0.33, sigma_t=2.5, sigma_rot_deg=0.15,
                                              sigma_scale=0.0015, blur_range=(0.8, 1.8), dtype="f32")
    return frames


@pytest.mark.parametrize("option", [1, 2, 3])
@pytest.mark.parametrize("acc,interp,motion", [(0, cv2.INTER_LINEAR, 0), (1, cv2.INTER_CUBIC, 3)])
def test_stack_with_upscale_after_align_matches_oracle(gpu, option, acc, interp, motion):
    """process_input_sequence with frame_upscale_after_align: the analytic map is up-scaled on the fly inside the fused kernel."""
    from serstacker_b200 import api
    frames = _sequence()
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE if acc else opl.ACC_AVERAGE, upscale_option=option)
    so.registration.motion_type = motion
    so.registration.interpolation = interp
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    avg_o, mask_o, _, _ = opl.run_stacking(frames, so)
    ro = api.registration_options(motion_type=motion, interpolation=interp, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=acc, max_batch=4, upscale_option=option,
                                                        upscale_stage=1))
    p.set_reference(frames[0])
    p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert avg_g.shape == avg_o.shape
    m = (mask_o > 0) & (mask_g > 0)
    rel = float(np.sqrt(((avg_g[m] - avg_o[m]) ** 2).sum()) / np.sqrt((avg_o[m] ** 2).sum()))
    print("  stack with upscale option %d (acc %d, interp %d, motion %d): %s, rel-L2 = %.3g, mask mismatch %.3g"
          % (option, acc, interp, motion, avg_g.shape, rel, (mask_o != mask_g).mean()))
    assert rel <= 1e-4
    assert (mask_o != mask_g).mean() < 1e-3


def test_stack_with_eccflow_and_upscale_matches_oracle(gpu):
    """What the "Planetary Disk" preset runs (c_image_stacking_pipeline.cc:321-327): eccflow + x1.5 after align."""
    from serstacker_b200 import api
    from test_gpu_eccflow import _turbulent_sequence, _flow_registration_options
    frames = _turbulent_sequence(320, 240, 4, seed=41)
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE, upscale_option=opl.UPSCALE_X15)
    so.registration = _flow_registration_options(3, 3, cv2.INTER_CUBIC)
    avg_o, mask_o, _, _ = opl.run_stacking(frames, so)
    ro = api.registration_options(motion_type=3, interpolation=2, enable_eccflow_registration=1, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=4, upscale_option=2, upscale_stage=1))
    p.set_reference(frames[0])
    p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert avg_g.shape == (360, 480) == avg_o.shape
    m = (mask_o > 0) & (mask_g > 0)
    rel = float(np.sqrt(((avg_g[m] - avg_o[m]) ** 2).sum()) / np.sqrt((avg_o[m] ** 2).sum()))
    print("  stack with eccflow + x1.5: rel-L2 = %.3g, mask mismatch %.3g" % (rel, (mask_o != mask_g).mean()))
    assert rel <= 1e-4 and (mask_o != mask_g).mean() < 1e-3


def test_stack_upscale_rejects(gpu):
    from serstacker_b200 import api
    ro = api.registration_options(motion_type=0)
    with pytest.raises(api.SskError):
        api.c_image_stacking_pipeline(api.stack_options(registration=ro, upscale_option=2, upscale_stage=2))      # before_align: not fused
    with pytest.raises(api.SskError):
        api.c_image_stacking_pipeline(api.stack_options(registration=ro, upscale_option=7, upscale_stage=1))
