"""GPU parity: accumulators and the fused per-frame loop (K5, P1) against the oracle pipeline."""
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


@pytest.mark.parametrize("dtype,cn", [(np.float32, 1), (np.float32, 3), (np.uint16, 1), (np.uint8, 3)])
@pytest.mark.parametrize("wmode", ["none", "mask", "weights"])
def test_weighted_average_add_matches_oracle(gpu, dtype, cn, wmode):
    from serstacker_b200 import api
    rng = np.random.default_rng(1)
    h, w = 37, 53
    o, g = oacc.WeightedAverage(), api.c_weigthed_average()
    bpp = {np.float32: 0, np.uint16: 16, np.uint8: 8}[dtype]
    for i in range(6):
        if dtype == np.float32:
            f = rng.random((h, w, cn)).astype(np.float32)
        else:
            f = rng.integers(0, np.iinfo(dtype).max, (h, w, cn)).astype(dtype)
        f = f.reshape(h, w) if cn == 1 else f
        wts = None
        if wmode == "mask":
            wts = ((rng.random((h, w)) > 0.3) * 255).astype(np.uint8)
        elif wmode == "weights":
            wts = (rng.random((h, w)) - 0.2).astype(np.float32)   # some non-positive weights are skipped
        o.add(opl.to_float_frame(f, bpp) if dtype != np.float32 else f, wts)
        g.add(f, wts, bpp=bpp)
    ao, mo = o.compute()
    ag, mg = g.compute()
    assert g.accumulated_frames() == 6
    assert np.array_equal(mo, mg)
    assert np.abs(ag - ao).max() <= 2e-6 * max(1.0, np.abs(ao).max())
    assert np.allclose(g.get_acc_counters(), o.get_acc_counters(), rtol=1e-6, atol=1e-6)


def test_accumulator_reinitialize_and_sum_form(gpu):
    from serstacker_b200 import api, capi
    rng = np.random.default_rng(2)
    a = rng.random((20, 30)).astype(np.float32)
    w = (rng.random((20, 30)) * 5).astype(np.float32)
    g = api.c_weigthed_average()
    g.reinitialize(a, w)
    capi.check(capi.lib.ssk_acc_to_sum_form(g._h))
    s, _ = g.compute()
    assert np.allclose(s, a * w, rtol=1e-6)
    capi.check(capi.lib.ssk_acc_from_sum_form(g._h, 3))
    b, _ = g.compute()
    assert np.allclose(b, a, rtol=1e-6, atol=1e-7)
    assert g.accumulated_frames() == 3


@pytest.mark.parametrize("colorid", [8, 9, 10, 11])
@pytest.mark.parametrize("mapped", [False, True])
def test_bayer_average_matches_oracle(gpu, colorid, mapped):
    from serstacker_b200 import api
    frames, shifts, bpp = synth.make_bayer_sequence(96, 64, 4, seed=3)
    o, g = oacc.BayerAverage(), api.c_bayer_average()
    o.set_bayer_pattern(colorid)
    g.set_bayer_pattern(colorid)
    rng = np.random.default_rng(4)
    for f, (tx, ty) in zip(frames, shifts):
        ff = opl.to_float_frame(f, bpp)
        mask = ((rng.random(f.shape) > 0.1) * 255).astype(np.uint8)
        if mapped:
            t = otf.TranslationTransform(-tx + 0.3, -ty - 0.4)
            rmap = t.create_remap((96, 64))
            o.set_remap(rmap)
            g.set_remap(rmap=rmap)
        o.add(ff, mask)
        g.add(f, mask, bpp=bpp)
    ao, mo = o.compute()
    ag, mg = g.compute()
    assert np.array_equal(mo, mg)
    assert np.abs(ag - ao).max() <= 1e-6


def _config1_like(n=10, size=(320, 240)):
    frames, mats, bpp = synth.make_planet_sequence(size[0], size[1], n, seed=1, radius=min(size) * 0.31, sigma_t=3.0, dtype="u16")
    return frames, bpp


@pytest.mark.parametrize("method", [oecc.ECC_ALIGN_FORWARD_ADDITIVE, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM])
@pytest.mark.parametrize("maxlevel", [0, -1])
def test_stack_average_translation_matches_oracle(gpu, method, maxlevel):
    """Config #1 shape: mono16 SER frames, translation ECC, LINEAR/REFLECT101 warp, average."""
    from serstacker_b200 import api
    frames, bpp = _config1_like()
    so = opl.StackingOptions()
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    so.registration.ecc.ecc_method = method
    so.registration.ecc.ecch_max_level = maxlevel
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_stacking([opl.to_float_frame(f, bpp) for f in frames], so, collect=rec)

    ro = api.registration_options(motion_type=0, ecc=dict(ecc_method=method, ecch_max_level=maxlevel))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=0, max_batch=4))
    p.set_reference(frames[0], bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == sum(r["ok"] for r in rec)
    for rg, r in zip(res, rec):
        assert rg["ok"] == r["ok"]
        assert map_diff_px(0, rg["params"], r["params"], (320, 240)) <= 1e-3
        assert rg["iterations"] == r["iterations"]
    assert np.array_equal(mask_g, mask_o)
    m = mask_o > 0
    assert rel_l2(avg_g, avg_o, m) <= 1e-4
    wg = p.accumulator().get_acc_counters()
    assert np.array_equal(wg, acc_o.weights)


@pytest.mark.parametrize("method", [oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM, oecc.ECC_ALIGN_FORWARD_ADDITIVE])
def test_stack_weighted_affine_cubic_matches_oracle(gpu, method):
    """Config #2 shape: mono 32F, affine ECCH with translation-first, CUBIC warp, sharpness-weighted average."""
    from serstacker_b200 import api
    frames, mats, _ = synth.make_planet_sequence(480, 270, 8, seed=2, radius=100, sigma_t=4.0, sigma_rot_deg=0.2,
                                                 sigma_scale=0.002, blur_range=(0.8, 2.5), dtype="f32")
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = cv2.INTER_CUBIC
    so.registration.ecc.ecc_method = method
    so.registration.ecc.ecch_max_level = -1
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_stacking(frames, so, collect=rec)

    ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=method, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=8))
    p.set_reference(frames[0])
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    if method == oecc.ECC_ALIGN_FORWARD_ADDITIVE:
        # forward-additive affine is the least stable solver: compare within the reference's own envelope
        rec2 = []
        with dot_noise():
            avg_n, mask_n, acc_n, _ = opl.run_stacking(frames, so, collect=rec2)
        for rg, r, r2 in zip(res, rec, rec2):
            assert rg["ok"] == r["ok"]
            env = map_diff_px(3, r2["params"], r["params"], (480, 270))
            assert map_diff_px(3, rg["params"], r["params"], (480, 270)) <= max(1e-3, 4 * env)
        m = (mask_o > 0) & (mask_g > 0)
        assert rel_l2(avg_g, avg_o, m) <= max(1e-4, 4 * rel_l2(avg_n, avg_o, m & (mask_n > 0)))
        return
    for rg, r in zip(res, rec):
        assert rg["ok"] == r["ok"]
        assert map_diff_px(3, rg["params"], r["params"], (480, 270)) <= 1e-3, (rg, r)
    assert np.array_equal(mask_g, mask_o)
    m = mask_o > 0
    assert rel_l2(avg_g, avg_o, m) <= 1e-4
    assert rel_l2(p.accumulator().get_acc_counters(), acc_o.weights, m) <= 1e-4


@pytest.mark.parametrize("acc", [0, 1])
def test_stack_lanczos4_matches_oracle(gpu, acc):
    """ECC_INTER_LANCZOS4 as registration_options.interpolation (ecc2.h:38; base_remap, c_frame_registration.cc:1265-1386):
    frame, weight map and validity mask (8-bit fixed-point Lanczos table) through the one-thread-per-pixel fused kernel."""
    from serstacker_b200 import api
    frames, mats, _ = synth.make_planet_sequence(320, 240, 5, seed=6, radius=80, sigma_t=3.0, sigma_rot_deg=0.2,
                                                 sigma_scale=0.002, blur_range=(0.8, 2.0), dtype="f32")
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE if acc else opl.ACC_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = cv2.INTER_LANCZOS4
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    avg_o, mask_o, acc_o, _ = opl.run_stacking(frames, so)
    ro = api.registration_options(motion_type=3, interpolation=cv2.INTER_LANCZOS4, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=acc, max_batch=5))
    p.set_reference(frames[0])
    p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert np.array_equal(mask_g, mask_o)
    m = mask_o > 0
    print("  stack LANCZOS4 (acc %d): rel-L2 = %.3g" % (acc, rel_l2(avg_g, avg_o, m)))
    assert rel_l2(avg_g, avg_o, m) <= 1e-4


def test_batched_equals_frame_by_frame(gpu):
    """Batching must not change the result: the accumulation order inside a batch is the frame order."""
    from serstacker_b200 import api
    frames, bpp = _config1_like(n=9)
    outs = []
    for mb in (1, 4, 16):
        ro = api.registration_options(motion_type=0, ecc=dict(ecc_method=oecc.ECC_ALIGN_FORWARD_ADDITIVE))
        p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=0, max_batch=mb))
        p.set_reference(frames[0], bpp=bpp)
        p.add_frames(frames, want_results=False)
        outs.append(p.compute()[0])
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_host_frames_pipelined_upload_equals_device_frames(gpu):
    """Host frames go through sub-chunks whose upload overlaps the previous sub-chunk's processing (two slot sets at
    max_batch 32): results - records and stack - must equal the unpipelined path, for a frame count that leaves a
    ragged tail."""
    from serstacker_b200 import api
    frames, bpp = _config1_like(n=45, size=(160, 120))
    outs, recs = [], []
    for mb in (1, 32):
        ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=3, ecch_max_level=-1))
        p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=mb))
        p.set_reference(frames[0], bpp=bpp)
        recs.append(p.add_frames(frames))
        outs.append(p.compute())
    assert [r["ok"] for r in recs[0]] == [r["ok"] for r in recs[1]]
    for a, b in zip(recs[0], recs[1]):
        assert np.array_equal(a["params"], b["params"]) and a["iterations"] == b["iterations"]
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][0], outs[1][0])


def test_streaming_submit_wait_equals_add_frames(gpu):
    """ssk_stack_submit / ssk_stack_wait: chunks in flight, results fetched one chunk late, same records and stack."""
    from serstacker_b200 import api
    frames, bpp = _config1_like(n=41, size=(160, 120))
    ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=3, ecch_max_level=-1))
    so = api.stack_options(registration=ro, accumulation_method=1, max_batch=8)
    a = api.c_image_stacking_pipeline(so)
    a.set_reference(frames[0], bpp=bpp)
    want = a.add_frames(frames)
    b = api.c_image_stacking_pipeline(so)
    b.set_reference(frames[0], bpp=bpp)
    got, prev = [], None
    for i in range(0, len(frames), 8):
        t = b.submit(frames[i:i + 8])
        if prev is not None:
            got += b.wait(prev)
        prev = t
    got += b.wait(prev)
    assert len(got) == len(want)
    for x, y in zip(got, want):
        assert x["ok"] == y["ok"] and np.array_equal(x["params"], y["params"]) and x["iterations"] == y["iterations"]
    (avg_a, mask_a), (avg_b, mask_b) = a.compute(), b.compute()
    assert np.array_equal(mask_a, mask_b) and np.array_equal(avg_a, avg_b)
    assert a.accumulated_frames() == b.accumulated_frames()
    from serstacker_b200 import capi
    assert capi.lib.ssk_stack_wait(b._h, 0, None, None, 0, None) != 0      # ticket 0 left the 4-chunk ring long ago


def _bayer_oracle(frames, bpp, colorid, interpolation=cv2.INTER_LINEAR, enable_registration=True):
    so = opl.StackingOptions(accumulation_method=opl.ACC_BAYER_AVERAGE, enable_registration=enable_registration)
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    so.registration.interpolation = interpolation
    rec = []
    avg, mask, acc, _ = opl.run_bayer_stacking(frames, bpp, so, colorid, collect=rec)
    return avg, mask, acc, rec


@pytest.mark.parametrize("colorid", [8, 9, 10, 11])
@pytest.mark.parametrize("interp", [cv2.INTER_LINEAR, cv2.INTER_CUBIC])
def test_stack_bayer_average_matches_oracle(gpu, colorid, interp):
    """The bayer_average form of the batched loop (c_image_stacking_pipeline.cc:1730-1752): raw Bayer frames in, the
    device demosaics them for the registration and gathers the raw samples through each frame's remap under the eroded
    remap mask.  Translation trajectories are bit-faithful; sums / counters are compared exactly."""
    from serstacker_b200 import api
    frames, shifts, bpp = synth.make_bayer_sequence(192, 128, 6, seed=5)
    avg_o, mask_o, acc_o, rec = _bayer_oracle(frames, bpp, colorid, interp)
    ro = api.registration_options(motion_type=0, interpolation=interp)
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=2, bayer_colorid=colorid, max_batch=4))
    p.set_reference(frames[0], bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == sum(r["ok"] for r in rec)
    assert [rg["ok"] for rg in res] == [r["ok"] for r in rec]
    dmax = max(map_diff_px(0, rg["params"], r["params"], (192, 128)) for rg, r in zip(res, rec))
    print("bayer stack colorid=%d interp=%d: max|dparam| = %.3g px" % (colorid, interp, dmax))
    assert dmax <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    if dmax == 0:      # identical maps: identical double-precision gathers
        assert np.array_equal(avg_g, avg_o)
        assert np.array_equal(p.accumulator().get_acc_counters(), acc_o.get_acc_counters())
    else:
        assert rel_l2(avg_g, avg_o, mask_o > 0) <= 1e-4


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_stack_bayer_average_other_depths(gpu, dtype):
    from serstacker_b200 import api
    frames, shifts, bpp = synth.make_bayer_sequence(160, 96, 4, seed=7)
    if dtype == "u8":
        frames, bpp = [(f >> 8).astype(np.uint8) for f in frames], 8
    else:
        frames, bpp = [(f.astype(np.float32) / np.float32(65536.0)) for f in frames], 32
    avg_o, mask_o, acc_o, rec = _bayer_oracle(frames, bpp, 8)
    p = api.c_image_stacking_pipeline(api.stack_options(registration=api.registration_options(motion_type=0), accumulation_method=2,
                                                        bayer_colorid=8, max_batch=8))
    p.set_reference(frames[0], bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert all(rg["ok"] == r["ok"] for rg, r in zip(res, rec))
    assert max(map_diff_px(0, rg["params"], r["params"], (160, 96)) for rg, r in zip(res, rec)) <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rel_l2(avg_g, avg_o, mask_o > 0) <= 1e-5


@pytest.mark.parametrize("dtype", ["u8", "u16", "f32"])
@pytest.mark.parametrize("size", [(192, 128), (190, 126), (66, 98)])
@pytest.mark.parametrize("colorid", [8, 11])
def test_stack_bayer_fused_ecc_image_is_bit_identical(gpu, dtype, size, colorid):
    """ecc.scale 0.5: the one-pass raw -> pyrDown(gray(debayer_nn2)) kernel (k_bayer_gray_pyrdown, aligned 4-sample loads or
    scalar loads by the row pitch) gives the registration the same ECC image as debayer_nn2 -> k_pyrdown: identical
    transforms, sums and counters.  SSK_NO_BAYER_PYRDOWN selects the three-kernel chain."""
    import os
    from serstacker_b200 import api
    w, h = size
    frames, _, bpp = synth.make_bayer_sequence(w, h, 5, seed=17)
    if dtype == "u8":
        frames, bpp = [(f >> 8).astype(np.uint8) for f in frames], 8
    elif dtype == "f32":
        frames, bpp = [(f.astype(np.float32) / np.float32(65536.0)) for f in frames], 32
    out = []
    for chain in (False, True):
        if chain:
            os.environ["SSK_NO_BAYER_PYRDOWN"] = "1"
        try:
            p = api.c_image_stacking_pipeline(api.stack_options(registration=api.registration_options(motion_type=0), accumulation_method=2,
                                                                bayer_colorid=colorid, max_batch=3))
            p.set_reference(frames[0], bpp=bpp)
            res = p.add_frames(frames)
            avg, mask = p.compute()
            out.append((res, avg, mask, p.accumulator().get_acc_counters()))
        finally:
            os.environ.pop("SSK_NO_BAYER_PYRDOWN", None)
    (ra, avg_a, mask_a, cnt_a), (rb, avg_b, mask_b, cnt_b) = out
    assert [r["ok"] for r in ra] == [r["ok"] for r in rb] and any(r["ok"] for r in ra[1:])
    assert all(np.array_equal(x["params"], y["params"]) for x, y in zip(ra, rb))
    assert np.array_equal(avg_a, avg_b) and np.array_equal(mask_a, mask_b) and np.array_equal(cnt_a, cnt_b)


def test_stack_bayer_average_without_registration(gpu):
    """enable_registration = false: empty remap, every raw sample goes to its own colour plane (c_frame_accumulation.cc:998-1010)."""
    from serstacker_b200 import api
    frames, _, bpp = synth.make_bayer_sequence(64, 48, 3, seed=9)
    avg_o, mask_o, acc_o, _ = _bayer_oracle(frames, bpp, 9, enable_registration=False)
    p = api.c_image_stacking_pipeline(api.stack_options(accumulation_method=2, bayer_colorid=9, enable_registration=0, max_batch=2))
    p.set_reference(frames[0], bpp=bpp)
    p.add_frames(frames, want_results=False)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == 3
    assert np.array_equal(mask_g, mask_o) and np.array_equal(avg_g, avg_o)


def test_stack_bayer_average_float_master_over_u16_frames(gpu):
    """A CV_32F BGR master frame (what the master-frame pass hands over) over 16-bit raw frames."""
    from serstacker_b200 import api
    from oracle import debayer as od
    frames, _, bpp = synth.make_bayer_sequence(128, 96, 4, seed=11)
    master = opl.to_float_frame(od.debayer_nn2(frames[0], 8), bpp)
    avg_o, mask_o, _, rec = _bayer_oracle(frames, bpp, 8)
    p = api.c_image_stacking_pipeline(api.stack_options(registration=api.registration_options(motion_type=0), accumulation_method=2,
                                                        bayer_colorid=8, max_batch=4))
    p.set_reference(master, bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert max(map_diff_px(0, rg["params"], r["params"], (128, 96)) for rg, r in zip(res, rec)) <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rel_l2(avg_g, avg_o, mask_o > 0) <= 1e-5


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_stack_is_independent_of_the_ecc_cluster_size(gpu, cluster):
    """The ECC kernel runs 8 / 4 / 2 CTAs per frame depending on the batch length (ssk_engine.cu: 2 from 400 frames on, what
    bench.py's 1024-frame chunks use).  Every cluster size must give the oracle's registration and stack (config #2 shape)."""
    import os
    from serstacker_b200 import api
    frames, mats, _ = synth.make_planet_sequence(480, 270, 6, seed=4, radius=100, sigma_t=4.0, sigma_rot_deg=0.2,
                                                 sigma_scale=0.002, blur_range=(0.8, 2.5), dtype="f32")
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = cv2.INTER_CUBIC
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_stacking(frames, so, collect=rec)
    os.environ["SSK_ECC_CLUSTER"] = str(cluster)          # read when the handle is created
    try:
        ro = api.registration_options(motion_type=3, interpolation=2, ecc=dict(ecc_method=3, ecch_max_level=-1))
        p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=8))
    finally:
        del os.environ["SSK_ECC_CLUSTER"]
    p.set_reference(frames[0])
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    worst = 0.0
    for rg, r in zip(res, rec):
        assert rg["ok"] == r["ok"]
        worst = max(worst, map_diff_px(3, rg["params"], r["params"], (480, 270)))
    m = mask_o > 0
    print("cluster %d: max |d map| = %.3g px, stack rel-L2 = %.3g" % (cluster, worst, rel_l2(avg_g, avg_o, m)))
    assert worst <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rel_l2(avg_g, avg_o, m) <= 1e-4


@pytest.mark.parametrize("interp", [cv2.INTER_CUBIC, cv2.INTER_LINEAR])
@pytest.mark.parametrize("dtype", ["f32", "u16"])
def test_stack_colour_frames_matches_oracle(gpu, interp, dtype):
    """Colour (BGR) frames through the loop: ECC on cvtColor(BGR2GRAY), every channel warped with the frame's map, one weight
    map per frame (c_image_stacking_pipeline.cc:1644-1714).  Exercises the multi-channel form of the fused kernel, whose
    interior pixels share the tap coefficients between the weight map and the channels."""
    from serstacker_b200 import api
    mono, mats, _ = synth.make_planet_sequence(400, 300, 6, seed=9, radius=110, sigma_t=3.0, sigma_rot_deg=0.15,
                                               sigma_scale=0.002, blur_range=(0.8, 2.0), dtype="f32")
    gains = np.array([0.85, 1.0, 0.7], np.float32)
    frames = [np.ascontiguousarray(f[..., None] * gains) for f in mono]
    bpp = 0
    if dtype == "u16":
        frames = [np.rint(f * 65535).astype(np.uint16) for f in frames]
        bpp = 16
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_AFFINE
    so.registration.interpolation = interp
    so.registration.ecc.ecc_method = oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM
    so.registration.ecc.ecch_max_level = -1
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_stacking([opl.to_float_frame(f, bpp) for f in frames], so, collect=rec)
    ro = api.registration_options(motion_type=3, interpolation={cv2.INTER_LINEAR: 1, cv2.INTER_CUBIC: 2}[interp],
                                  ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1, max_batch=8))
    p.set_reference(frames[0], bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert avg_g.shape == avg_o.shape == (300, 400, 3)
    worst = 0.0
    for rg, r in zip(res, rec):
        assert rg["ok"] == r["ok"]
        worst = max(worst, map_diff_px(3, rg["params"], r["params"], (400, 300)))
    m = mask_o > 0
    rl = rel_l2(avg_g, avg_o, m)
    print("colour stack interp=%d %s: max |d map| = %.3g px, rel-L2 = %.3g" % (interp, dtype, worst, rl))
    assert worst <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rl <= 1e-4
