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


def test_config3_full_size_chain_through_the_stack_loop(gpu):
    """Config #3 at 4096x3000 RGGB 16-bit through ssk_stack (accumulation_method = bayer_average): device debayer_nn2 ->
    gray ECC translation registration at scale 0.5 -> eroded remap mask -> Bayer gather of the raw samples
    (c_image_stacking_pipeline.cc:1358-1862 with c_bayer_average, :1730-1752)."""
    from serstacker_b200 import api
    frames, shifts, bpp = synth.make_bayer_sequence(4096, 3000, 4, seed=3)
    so = opl.StackingOptions(accumulation_method=opl.ACC_BAYER_AVERAGE)
    so.registration.motion_type = otf.IMAGE_MOTION_TRANSLATION
    rec = []
    avg_o, mask_o, acc_o, _ = opl.run_bayer_stacking(frames, bpp, so, 8, collect=rec)
    p = api.c_image_stacking_pipeline(api.stack_options(registration=api.registration_options(motion_type=0), accumulation_method=2,
                                                        bayer_colorid=8, max_batch=4))
    p.set_reference(frames[0], bpp=bpp)
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == sum(r["ok"] for r in rec) == len(frames)
    dmax = max(map_diff_px(0, rg["params"], r["params"], (4096, 3000)) for rg, r in zip(res, rec))
    m = (mask_o > 0) & (mask_g > 0)
    rl2 = rel_l2(avg_g, avg_o, m)
    print("config #3 full size: max|dparam| = %.3g px, stack rel-L2 = %.3g, mask mismatches = %d" % (dmax, rl2, int((mask_o != mask_g).sum())))
    assert dmax <= 1e-3
    assert np.array_equal(mask_g, mask_o)
    assert rl2 <= 1e-4
    assert rel_l2(p.accumulator().get_acc_counters(), acc_o.get_acc_counters(), m) <= 1e-5


def test_config4_full_size_jdr_align_derotate_average(gpu):
    """Config #4 at 2048x2048 (6 frames): c_jdr_pipeline's per-frame body - preproc_align_and_remap (c_ecch translation
    align with the c_ecch_options defaults, create_remap, cv::remap LINEAR / REPLICATE; c_jdr_pipeline.cc:546-590) and
    derotate + lpg-weighted blend (c_jdr_pipeline.cc:1184-1236).  Every frame restarts from the identity (SURVEY 8e:
    the reference's warm start is a sequential dependency; oracle and device run the same per-frame reset)."""
    import math
    from serstacker_b200 import api
    from oracle import derotation as od
    from test_gpu_derotation import _jovian_frame
    size, center = (2048, 2048), (1021.4, 1030.8)
    A = 700.0
    axes = (A, A * 0.93512560845968779724, A)          # c_jovian_ellipse_detector.h:67
    target = (0.3, math.radians(3.0), math.radians(15.0))
    period = 35740.632                                   # c_jovian_derotation_remap.h:19
    wts = 190.0                                          # c_jdr_pipeline_stack_options::wts
    times = [-150.0, -90.0, -30.0, 0.0, 60.0, 120.0]
    rng = np.random.default_rng(4)
    lpg_opts = dict(k=2.0, p=2.0, dscale=2, uscale=6)    # c_lpg_options defaults
    Rt = od.build_ellipsoid_rotation(*target)
    master = _jovian_frame(size, center, axes, target, seed=100)

    ot = otf.create_image_transform(0)
    oe = oecc.EccH(ot)                                   # c_ecch_options defaults: IC-LM, epsx 1e-5, maxlevel 0
    oe.set_reference_image(master, None)
    gt = api.create_image_transform(0)
    ge = api.c_ecch(gt)
    ge.set_reference_image(master)
    o, g = oacc.WeightedAverage(), api.c_weigthed_average()
    dmax = 0.0
    for i, dt in enumerate(times):
        dl = -2 * math.pi * dt / period                  # compute_derotation_for_time(-dt): rotation since the master
        pose = (target[0] + dl, target[1], target[2])
        shift = (0.0, 0.0) if dt == 0 else rng.normal(0.0, 2.5, 2)
        frame = _jovian_frame(size, (center[0] + shift[0], center[1] + shift[1]), axes, pose, seed=i)
        wscale = 1.0 / (1.0 + abs(dt) / wts)
        # oracle
        ot.reset()
        oe.align(frame, None)
        fo = cv2.remap(frame, oe.create_remap(), None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        od.jdr_derotate_and_add(o, fo, None, size, center, axes, target, dl, wscale, is_master=(dt == 0), lpg_opts=lpg_opts)
        # device
        gt.set_parameters(np.zeros(2, np.float32))
        ge.align(frame)
        dmax = max(dmax, map_diff_px(0, gt.parameters(), ot.parameters(), size))
        fg, _ = api.remap(gt, None, frame, interpolation=cv2.INTER_LINEAR, border_mode=cv2.BORDER_REPLICATE)
        _, _, _, ebox, cbox = od.compute_derotation_for_angle(size, center, axes, target, dl, wscale)
        Rc = od.build_ellipsoid_rotation(*pose)
        api.jdr_derotate_and_add(g, fg, None, center, axes, Rc, Rt, float(ebox[2]), cbox, wscale, dt == 0,
                                 enable_weighted_average=True, lpg_k=2.0, lpg_p=2.0, lpg_dscale=2, lpg_uscale=6)
    ao, mo = o.compute()
    ag, mg = g.compute()
    m = (mo > 0) & (mg > 0)
    rl2 = rel_l2(ag, ao, m)
    print("config #4 full size: max|dparam| = %.3g px, stack rel-L2 = %.3g, mask mismatch = %.3g" % (dmax, rl2, float(np.mean(mo != mg))))
    assert g.accumulated_frames() == len(times)
    assert dmax <= 1e-3
    assert np.mean(mo != mg) < 1e-4
    assert rl2 <= 1e-4
    assert rel_l2(g.get_acc_counters(), o.weights, m) <= 1e-4


def test_config5_full_size_focus_stack(gpu):
    """Config #5 at 2448x2048 RGB 32F (4 frames): W = lpg(k=6, p=2, dscale=0, uscale=0) = (6 lap^2 + grad^2)^2 scaled as
    lpg.cc:223-290, GaussianBlur(sigma = 1), c_weigthed_average::add(frame, W) - no registration."""
    from serstacker_b200 import api
    from oracle import weights as ow
    W5, H5 = 2448, 2048
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:H5, 0:W5].astype(np.float32)
    scene = np.zeros((H5, W5, 3), np.float32)
    for c in range(3):
        tex = rng.random((H5 // 8, W5 // 8)).astype(np.float32)
        tex = cv2.resize(tex, (W5, H5), interpolation=cv2.INTER_CUBIC)
        fine = rng.random((H5, W5)).astype(np.float32)
        scene[..., c] = np.clip(0.35 + 0.4 * (tex - 0.5) + 0.25 * (cv2.GaussianBlur(fine, (0, 0), 1.0) - 0.5) * 4, 0, 1)
    depth = (xx / W5 + 0.5 * yy / H5) / 1.5             # tilted depth map in [0, 1]
    o, g = oacc.WeightedAverage(), api.c_weigthed_average()
    for i, focus in enumerate((0.15, 0.4, 0.65, 0.9)):
        # depth-dependent defocus: blend of three blur levels by |depth - focus|
        d = np.clip(np.abs(depth - focus) * 3.0, 0, 1)[..., None]
        b1, b2 = cv2.GaussianBlur(scene, (0, 0), 1.5), cv2.GaussianBlur(scene, (0, 0), 4.0)
        frame = np.where(d < 0.5, scene * (1 - 2 * d) + b1 * (2 * d), b1 * (2 - 2 * d) + b2 * (2 * d - 1)).astype(np.float32)
        frame = np.clip(frame + rng.normal(0, 0.002, frame.shape).astype(np.float32), 0, 1)
        wo = cv2.GaussianBlur(ow.lpg(frame, k=6.0, p=2.0, dscale=0, uscale=0), (0, 0), 1, None, 1, cv2.BORDER_REPLICATE)
        o.add(frame, wo)
        wg = api.gaussian_blur(api.lpg(frame, k=6.0, p=2.0, dscale=0, uscale=0), 1.0)
        assert rel_l2(wg, wo) <= 1e-5
        g.add(frame, wg)
    ao, mo = o.compute()
    ag, mg = g.compute()
    rl2 = rel_l2(ag, ao, mo > 0)
    print("config #5 full size: stack rel-L2 = %.3g, weights rel-L2 = %.3g" % (rl2, rel_l2(g.get_acc_counters(), o.weights)))
    assert np.array_equal(mo, mg)
    assert rl2 <= 1e-4
    assert rel_l2(g.get_acc_counters(), o.weights) <= 1e-4
