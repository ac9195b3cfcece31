"""GPU parity: average_pyramid_inpaint (core/proc/inpaint/average_pyramid_inpaint.cc:97-127), the finishing step of the
stacking pass (c_image_stacking_pipeline.cc:763-767), against oracle/inpaint.py.  Bit-exact."""
import numpy as np
import pytest

from oracle import inpaint as oip
from oracle import pipeline as opl
from oracle import transforms as otf
from serstacker_b200 import synth
from test_inpaint_oracle import holes_mask

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.mark.parametrize("shape", [(2, 2), (5, 7), (33, 47), (64, 96), (135, 240), (90, 61, 3), (77, 50, 2), (40, 41, 4),
                                   (1, 9), (9, 1), (540, 960)])
@pytest.mark.parametrize("fill", [0.02, 0.5, 0.9])
def test_inpaint_matches_oracle(gpu, shape, fill):
    from serstacker_b200 import api
    rng = np.random.default_rng(shape[0] * 131 + int(fill * 100))
    src = rng.random(shape, dtype=f32)
    mask = holes_mask(rng, shape[0], shape[1], fill) if min(shape[:2]) > 1 else \
        (rng.random(shape[:2]) < fill).astype(np.uint8) * 255
    if mask.all():
        mask.flat[0] = 0
    want, wmask = oip.average_pyramid_inpaint(src, mask)
    got, gmask = api.average_pyramid_inpaint(src, mask)
    assert np.array_equal(gmask, wmask)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("max_levels", [0, 1, 2, 3, 100])
def test_inpaint_max_levels(gpu, max_levels):
    from serstacker_b200 import api
    rng = np.random.default_rng(5)
    src = rng.random((80, 120), dtype=f32)
    mask = holes_mask(rng, 80, 120, 0.6)
    want, wmask = oip.average_pyramid_inpaint(src, mask, max_levels)
    got, gmask = api.average_pyramid_inpaint(src, mask, max_levels)
    assert np.array_equal(gmask, wmask)
    assert np.array_equal(got, want)
    assert np.array_equal(got[mask > 0], src[mask > 0])


def test_inpaint_full_none_and_empty_masks(gpu):
    from serstacker_b200 import api
    rng = np.random.default_rng(6)
    src = rng.random((31, 45, 3), dtype=f32)
    full = np.full((31, 45), 200, np.uint8)           # no holes: copies, the mask values are kept
    got, gmask = api.average_pyramid_inpaint(src, full)
    assert np.array_equal(got, src) and np.array_equal(gmask, full)
    got, gmask = api.average_pyramid_inpaint(src, None)
    assert np.array_equal(got, src) and gmask is None
    none = np.zeros((31, 45), np.uint8)
    want, wmask = oip.average_pyramid_inpaint(src, none)
    got, gmask = api.average_pyramid_inpaint(src, none)
    assert np.array_equal(got, want) and np.array_equal(gmask, wmask)


def test_inpaint_non_binary_mask_and_wide_dynamic_range(gpu):
    from serstacker_b200 import api
    rng = np.random.default_rng(7)
    src = rng.random((70, 90), dtype=f32) * (10.0 ** rng.integers(-6, 1, size=(70, 90))).astype(f32)
    mask = holes_mask(rng, 70, 90, 0.5)
    mask[mask > 0] = rng.integers(1, 256, size=int((mask > 0).sum())).astype(np.uint8)
    want, wmask = oip.average_pyramid_inpaint(src, mask)
    got, gmask = api.average_pyramid_inpaint(src, mask)
    assert np.array_equal(gmask, wmask)
    assert np.array_equal(got, want)


def test_inpaint_rejects_bad_arguments(gpu):
    from serstacker_b200 import api
    src = np.zeros((8, 8), f32)
    with pytest.raises(Exception):
        api.average_pyramid_inpaint(src, np.zeros((8, 9), np.uint8))


def test_stack_compute_inpainted_matches_oracle(gpu):
    """The end of a run (c_image_stacking_pipeline.cc:742-767): compute() then average_pyramid_inpaint(.., 100), for
    shifted frames that leave part of the reference canvas uncovered."""
    from serstacker_b200 import api
    frames, _, bpp = synth.make_planet_sequence(320, 240, 3, seed=4, radius=70, sigma_t=6.0, dtype="u16")
    so = opl.StackingOptions()
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    fl = [opl.to_float_frame(f, bpp) for f in frames]
    avg_o, mask_o, acc_o, _ = opl.run_stacking(fl[1:2], so, reference=fl[0])

    ro = api.registration_options(motion_type=0)
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=0, max_batch=4))
    p.set_reference(frames[0], bpp=bpp)
    p.add_frames(frames[1:2])
    avg_g, mask_g = p.compute()
    assert np.array_equal(mask_g, mask_o)
    assert (mask_g == 0).sum() > 1000         # the uncovered strip
    # the device chain on the device's own average is exactly the oracle's inpaint of that average
    want, wmask = oip.average_pyramid_inpaint(avg_g, mask_g, 100)
    got, gmask = p.compute(inpaint_max_levels=100)
    assert np.array_equal(gmask, wmask)
    assert np.array_equal(got, want)
    assert gmask.min() == 255
    # and end to end against the oracle pipeline
    want_o, _ = oip.average_pyramid_inpaint(avg_o, mask_o, 100)
    assert np.abs(got - want_o).max() <= 1e-5 * max(1.0, float(np.abs(want_o).max()))
    got2, gmask2 = p.accumulator().compute_inpainted(1.0, 100)
    assert np.array_equal(got2, got) and np.array_equal(gmask2, gmask)


@pytest.mark.parametrize("cn", [1, 3])
def test_accumulator_compute_inpainted_matches_oracle(gpu, cn):
    from serstacker_b200 import api
    rng = np.random.default_rng(11)
    h, w = 120, 160
    g = api.c_weigthed_average()
    holes = holes_mask(rng, h, w, 0.8, boxes=3)
    for i in range(4):
        f = rng.random((h, w, cn), dtype=f32)
        f = f.reshape(h, w) if cn == 1 else f
        wts = (rng.random((h, w), dtype=f32) + f32(0.1)) * (holes > 0)
        g.add(f, wts.astype(f32))
    avg, mask = g.compute()
    assert np.array_equal(mask, holes)
    want, wmask = oip.average_pyramid_inpaint(avg, mask, 100)
    got, gmask = g.compute_inpainted(1.0, 100)
    assert np.array_equal(gmask, wmask)
    assert np.array_equal(got, want)


def test_stacking_pass_with_sharpened_reference_and_inpaint_matches_oracle(gpu):
    """Both neighbours of the per-frame loop in one run: unsharp_mask of the master frame (sigma 1, alpha 0.8:
    c_image_stacking_pipeline.cc:1302-1306), registration + weighted stacking, average_pyramid_inpaint of the result."""
    from serstacker_b200 import api
    from oracle import ecc as oecc
    from helpers import rel_l2
    import cv2
    frames, _, bpp = synth.make_planet_sequence(320, 240, 7, seed=9, radius=70, sigma_t=5.0, dtype="f32")
    master = np.mean(np.stack(frames[:3]), axis=0).astype(f32)       # any float master frame
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = cv2.INTER_CUBIC
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    rec = []
    avg_o, mask_o, ref_o = opl.run_stacking_pass(frames[3:], so, master, 1.0, 0.8, 100, collect=rec)

    ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=4))
    avg_g, mask_g, res = p.run_stacking_pass(frames[3:], master, bpp=bpp, unsharp_sigma=1.0, unsharp_alpha=0.8)
    assert p.accumulated_frames() == sum(r["ok"] for r in rec)
    for rg, r in zip(res, rec):
        assert rg["ok"] == r["ok"]
        assert np.abs(rg["params"] - r["params"]).max() <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rel_l2(avg_g, avg_o, mask_o > 0) <= 1e-4
