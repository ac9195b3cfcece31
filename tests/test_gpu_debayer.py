"""GPU parity: debayer_nn2 (core/io/debayer.cc:827-1195) against oracle/debayer.py.  Bit-exact for every depth."""
import numpy as np
import pytest

from oracle import debayer as od

pytestmark = pytest.mark.gpu


def _raw(rng, shape, dtype):
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    return rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("colorid", [8, 9, 10, 11])
@pytest.mark.parametrize("shape", [(2, 2), (2, 6), (6, 2), (10, 14), (250, 334), (480, 640), (2, 4), (6, 8), (252, 336), (4, 16), (6, 48)])
def test_debayer_nn2_matches_oracle(gpu, dtype, colorid, shape):
    from serstacker_b200 import api
    rng = np.random.default_rng(shape[0] + colorid)
    raw = _raw(rng, shape, dtype)
    assert np.array_equal(api.debayer_nn2(raw, colorid), od.debayer_nn2(raw, colorid))


def test_debayer_nn2_config3_size_and_saturated_values(gpu):
    """Config #3's frame: 4096x3000 RGGB16, including full-scale samples (the integer averages must not wrap)."""
    from serstacker_b200 import api
    rng = np.random.default_rng(0)
    raw = _raw(rng, (3000, 4096), np.uint16)
    raw[100:200, 100:300] = 65535
    assert np.array_equal(api.debayer_nn2(raw, 8), od.debayer_nn2(raw, 8))


def test_debayer_nn2_rejects_uneven_sizes_and_unknown_patterns(gpu):
    from serstacker_b200 import api
    with pytest.raises(Exception):
        api.debayer_nn2(np.zeros((5, 6), np.uint16), 8)
    with pytest.raises(Exception):
        api.debayer_nn2(np.zeros((6, 6), np.uint16), 3)


def test_config3_chain_debayer_register_bayer_average_matches_oracle(gpu):
    """Config #3's per-frame chain with the demosaic on the device: raw RGGB16 -> debayer_nn2 -> gray ECC registration ->
    remap validity mask -> Bayer accumulation of the raw samples through the frame's remap
    (c_image_stacking_pipeline.cc:1358-1862 with accumulation_method = bayer_average)."""
    from serstacker_b200 import api, synth
    from oracle import pipeline as opl, transforms as otf, accumulation as oacc
    from oracle.registration import FrameRegistration
    frames, shifts, bpp = synth.make_bayer_sequence(192, 128, 5, seed=5)
    so = opl.StackingOptions(accumulation_method=opl.ACC_BAYER_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    oreg = FrameRegistration(so.registration)
    oreg.setup_reference_frame(opl.to_float_frame(od.debayer_nn2(frames[0], 8), bpp), None)
    oa = oacc.BayerAverage()
    oa.set_bayer_pattern(8)
    oparams = []
    for f in frames:
        ok = opl.process_frame(oreg, oa, so, opl.to_float_frame(od.debayer_nn2(f, 8), bpp), None,
                               raw_bayer=opl.to_float_frame(f, bpp))
        assert ok
        oparams.append(oreg.image_transform.clone_parameters())
    avg_o, mask_o = oa.compute()

    greg = api.c_frame_registration(api.registration_options(motion_type=0))
    ga = api.c_bayer_average()
    ga.set_bayer_pattern(8)
    greg.setup_reference_frame(api.debayer_nn2(frames[0], 8), bpp=bpp)
    for f, po in zip(frames, oparams):
        assert greg.register_frame(api.debayer_nn2(f, 8), bpp=bpp)
        assert np.abs(greg.image_transform_parameters() - po).max() <= 1e-3
        rmap = greg.current_remap()
        _, mask = greg.custom_remap(rmap, None, None, want_mask=True)
        ga.set_remap(rmap=rmap)
        ga.add(f, mask, bpp=bpp)
    avg_g, mask_g = ga.compute()
    assert ga.accumulated_frames() == len(frames)
    assert np.mean(mask_g == mask_o) >= 0.999
    m = (mask_g > 0) & (mask_o > 0)
    assert np.abs(avg_g - avg_o)[m].max() <= 2e-4
