"""GPU parity at the FULL frame sizes of BASELINE.json's configurations (a few frames each, so that the oracle still
finishes in seconds), plus size-independent properties of the config #2 path at 1920x1080: identical frames stack to
themselves, and the result does not depend on how the sequence is cut into batches."""
import numpy as np
import cv2
import pytest

from oracle import accumulation as oacc
from oracle import ecc as oecc
from oracle import pipeline as opl
from oracle import transforms as otf
from serstacker_b200 import synth
from helpers import map_diff_px, rel_l2, dot_noise

pytestmark = pytest.mark.gpu

W2, H2 = 1920, 1080


@pytest.fixture(scope="module")
def frames_1080p():
    frames, _, _ = synth.make_planet_sequence(W2, H2, 5, seed=2, radius=400, sigma_t=4.0, sigma_rot_deg=0.2,
                                              sigma_scale=0.002, blur_range=(0.8, 2.5), dtype="f32")
    return frames


def _config2_pipeline(max_batch):
    from serstacker_b200 import api
    ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=3, ecch_max_level=-1))
    return api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=max_batch))


def test_config2_full_size_matches_oracle(gpu, frames_1080p):
    """Config #2 at 1920x1080: affine ECCH (IC-LM, translation first, full pyramid), CUBIC warp, weighted average.
    north_star: parameters within 1e-3 px, stack within 1e-4 relative L2."""
    frames = frames_1080p
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = cv2.INTER_CUBIC
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_stacking(frames, so, collect=rec)
    rec_n = []
    with dot_noise():          # the oracle's own sensitivity envelope at this size (helpers.py)
        opl.run_stacking(frames, so, collect=rec_n)

    p = _config2_pipeline(max_batch=8)
    p.set_reference(frames[0])
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == sum(r["ok"] for r in rec) == len(frames)
    for rg, r, rn in zip(res, rec, rec_n):
        assert rg["ok"] == r["ok"]
        env = map_diff_px(3, rn["params"], r["params"], (W2, H2))
        assert map_diff_px(3, rg["params"], r["params"], (W2, H2)) <= max(1e-3, 4 * env), (rg, r, env)
    m = (mask_o > 0) & (mask_g > 0)
    assert (mask_o > 0).sum() - m.sum() <= 64 and (mask_g > 0).sum() - m.sum() <= 64
    assert rel_l2(avg_g, avg_o, m) <= 1e-4
    assert rel_l2(p.accumulator().get_acc_counters(), acc_o.weights, m) <= 1e-4


def test_identical_frames_stack_to_themselves_full_size(gpu, frames_1080p):
    """Idempotence: a sequence of copies of the reference registers to the identity and averages to the frame."""
    f = frames_1080p[0]
    p = _config2_pipeline(max_batch=4)
    p.set_reference(f)
    res = p.add_frames([f, f, f, f, f, f])
    avg, mask = p.compute()
    ident = np.array([1, 0, 0, 0, 1, 0], np.float32)
    for r in res:
        assert r["ok"]
        assert np.abs(r["params"][:6] - ident).max() <= 1e-6, r
    m = mask > 0
    assert m[8:-8, 8:-8].all()                      # only the ring the bicubic taps / 5x5 erosion cannot cover is masked
    assert np.abs(avg[m] - f[m]).max() <= 2e-6


def test_batching_does_not_change_the_stack_full_size(gpu, frames_1080p):
    """The sequence cut into batches of 5, 2 and 1 frames gives bit-identical registrations and stacks."""
    outs = []
    for mb in (5, 2, 1):
        p = _config2_pipeline(max_batch=mb)
        p.set_reference(frames_1080p[0])
        res = p.add_frames(frames_1080p)
        avg, mask = p.compute()
        outs.append((res, avg, mask))
    for res, avg, mask in outs[1:]:
        for a, b in zip(res, outs[0][0]):
            assert a["ok"] == b["ok"] and np.array_equal(a["params"], b["params"]) and a["iterations"] == b["iterations"]
        assert np.array_equal(mask, outs[0][2])
        assert np.array_equal(avg, outs[0][1])


@pytest.mark.parametrize("method", [oecc.ECC_ALIGN_FORWARD_ADDITIVE, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM])
def test_config1_full_size_matches_oracle(gpu, method):
    """Config #1 at 640x480 mono16: translation ECC (single level, the default), LINEAR / REFLECT101 warp, average."""
    from serstacker_b200 import api
    frames, _, bpp = synth.make_planet_sequence(640, 480, 16, seed=1, radius=150, sigma_t=3.0, dtype="u16")
    so = opl.StackingOptions()
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    so.registration.ecc.ecc_method = method
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_stacking([opl.to_float_frame(f, bpp) for f in frames], so, collect=rec)
    ro = api.registration_options(motion_type=0, ecc=dict(ecc_method=method))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=0, max_batch=16))
    p.set_reference(frames[0], bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == sum(r["ok"] for r in rec)
    for rg, r in zip(res, rec):
        assert rg["ok"] == r["ok"]
        assert map_diff_px(0, rg["params"], r["params"], (640, 480)) <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rel_l2(avg_g, avg_o, mask_o > 0) <= 1e-4
    assert np.array_equal(p.accumulator().get_acc_counters(), acc_o.weights)


def test_config3_full_size_bayer_average_matches_oracle(gpu):
    """Config #3 at 4096x3000 RGGB 16-bit: Bayer-pattern accumulation through a translation remap (two frames)."""
    from serstacker_b200 import api
    frames, shifts, bpp = synth.make_bayer_sequence(4096, 3000, 2, seed=3)
    o, g = oacc.BayerAverage(), api.c_bayer_average()
    o.set_bayer_pattern(8)
    g.set_bayer_pattern(8)
    for f, (tx, ty) in zip(frames, shifts):
        rmap = otf.TranslationTransform(-tx + 0.3, -ty - 0.4).create_remap((4096, 3000))
        o.set_remap(rmap)
        g.set_remap(rmap=rmap)
        o.add(opl.to_float_frame(f, bpp), None)
        g.add(f, None, bpp=bpp)
    ao, mo = o.compute()
    ag, mg = g.compute()
    assert np.array_equal(mo, mg)
    assert np.abs(ag - ao).max() <= 1e-6
