"""GPU parity: c_eccflow (ecc2.cc:2220-2865) through the C ABI against oracle/eccflow.py (the reference restated over cv2).

Tolerances.  The device forms the INTER_AREA sums column-first and fuses multiply-adds where OpenCV forms them row-first, so
one iteration of one level agrees with the oracle to float rounding (measured: mean 1e-7 px, max 3e-5 px = one ulp of the
map coordinate).  The coarse-to-fine recursion itself amplifies rounding: the flow feeds cv::remap, which quantises
coordinates to 1/32 px, every level multiplies the flow by 4/3 and the update multiplier 1.5 over-relaxes.  The oracle
re-run on the same frame scaled by (1 + 1e-7 noise) - one float ulp - moves its own answer by ~3e-4 px on average and
~1.5e-3 px at most over the 18 levels of a 480 x 270 image (tools/flow_debug.py), so the reference's result is only
defined up to that envelope (it depends on the SIMD width of the OpenCV build).  The contract asserted here is therefore
the measured envelope, like the ECC parity tests do: 99.9 % of the map within max(1e-3 px, 3 x the oracle's own 99.9 %
spread), every pixel within max(2e-3 px, 4 x its maximum spread); each test prints what it measured."""
import numpy as np
import cv2
import pytest

from oracle import eccflow as oef

pytestmark = pytest.mark.gpu
f32 = np.float32


def _scene(h, w, seed, amp=40.0, smooth=25.0, noise=0.002):
    """Textured reference + a copy warped by a smooth random displacement field (atmospheric-turbulence-like)."""
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.random((h + 40, w + 40)).astype(f32), (0, 0), 2.0)
    base = (base - base.min()) / (base.max() - base.min())
    ref = base[20:20 + h, 20:20 + w].copy()
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    du = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), smooth) * amp
    dv = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), smooth) * amp
    cur = cv2.remap(base, xx + 20 + du, yy + 20 + dv, cv2.INTER_CUBIC)
    cur = (cur + rng.standard_normal(cur.shape).astype(f32) * noise).astype(f32)
    return ref, cur, np.stack([xx, yy], -1)


def _gpu_options(o):
    from serstacker_b200 import api
    return api.eccflow_options(update_multiplier=o.update_multiplier, scale_factor=o.scale_factor, noise_level=o.noise_level,
                               max_iterations=o.max_iterations, support_scale=o.support_scale, min_image_size=o.min_image_size,
                               max_pyramid_level=o.max_pyramid_level, downscale_method=o.downscale)


def _check(name, got, want, fo, cur, rmap0=None, mask=None):
    """Compares the device map with the oracle's against the oracle's own one-ulp sensitivity envelope."""
    rng = np.random.default_rng(12345)
    cur2 = (cur * (1 + rng.standard_normal(cur.shape).astype(f32) * f32(1e-7))).astype(f32)
    env = np.abs(fo.compute(cur2, rmap0, mask) - want).max(axis=-1)
    d = np.abs(got - want).max(axis=-1)
    p999, e999 = float(np.quantile(d, 0.999)), float(np.quantile(env, 0.999))
    print("  eccflow %s: max |d map| = %.3g px, 99.9%% = %.3g px, mean = %.3g px  (oracle one-ulp envelope: max %.3g, 99.9%% %.3g, mean %.3g)"
          % (name, d.max(), p999, d.mean(), env.max(), e999, env.mean()))
    assert p999 <= max(1e-3, 3 * e999), (p999, e999)
    assert d.max() <= max(2e-3, 4 * float(env.max())), (d.max(), env.max())
    assert d.mean() <= max(1e-4, 3 * float(env.mean())), (d.mean(), env.mean())


@pytest.mark.parametrize("shape", [(270, 480), (203, 331), (64, 300)])
@pytest.mark.parametrize("method", [oef.DOWNSCALE_RECURSIVE_RESIZE, oef.DOWNSCALE_FULL_RESIZE, oef.DOWNSCALE_PYRAMID])
def test_eccflow_reference_pyramid_matches_oracle(gpu, shape, method):
    """set_reference_image: level sizes (incl. the big-aspect-ratio rule), images, gradients and the D field."""
    from serstacker_b200 import api
    ref, _, _ = _scene(shape[0], shape[1], seed=shape[1])
    o = oef.registration_options(downscale=method)
    want = oef.EccFlow(o)
    want.set_reference_image(ref)
    got = api.c_eccflow(_gpu_options(o))
    got.set_reference_image(ref)
    assert got.num_levels() == len(want.pyramid)
    for l, e in enumerate(want.pyramid):
        w, h, gw, gh = got.level_size(l)
        assert (h, w) == e.reference_image.shape and (gh, gw) == e.D.shape[:2]
        assert np.abs(got.pyramid_image(0, l) - e.reference_image).max() <= 2e-7
        assert np.abs(got.pyramid_image(2, l) - e.Ix).max() <= 1e-6
        assert np.abs(got.pyramid_image(3, l) - e.Iy).max() <= 1e-6
        D = got.pyramid_image(4, l)
        assert np.allclose(D[..., :3], e.D[..., :3], rtol=1e-4, atol=1e-9)
        # 1 / det amplifies the rounding of a cancelling determinant: compare where the tensor is well conditioned
        det = np.abs(e.D[..., 0] * e.D[..., 2] - e.D[..., 1] ** 2)
        good = det > 1e-3 * e.D[..., 0] * e.D[..., 2]
        if good.any():
            assert np.allclose(D[..., 3][good], e.D[..., 3][good], rtol=2e-3)


def test_eccflow_single_iteration_is_rounding_exact(gpu):
    """One iteration of one level (no recursion to amplify anything): the device equals the oracle to one ulp of the map."""
    from serstacker_b200 import api
    ref, cur, ident = _scene(270, 480, 1)
    o = oef.registration_options(max_pyramid_level=0, max_iterations=1)
    fo = oef.EccFlow(o)
    fo.set_reference_image(ref)
    want = fo.compute(cur, ident)
    fg = api.c_eccflow(_gpu_options(o))
    fg.set_reference_image(ref)
    got = fg.compute(cur, ident)
    d = np.abs(got - want).max(axis=-1)
    print("  eccflow one level, one iteration: max |d map| = %.3g px, mean %.3g px" % (d.max(), d.mean()))
    assert d.max() <= 6.2e-5 and d.mean() <= 1e-6


@pytest.mark.parametrize("shape,seed", [((270, 480), 1), ((203, 331), 2), ((540, 960), 3)])
@pytest.mark.parametrize("initial", ["empty", "identity", "affine"])
def test_eccflow_compute_matches_oracle(gpu, shape, seed, initial):
    from serstacker_b200 import api
    h, w = shape
    ref, cur, ident = _scene(h, w, seed)
    o = oef.registration_options()
    rmap0 = None
    if initial == "identity":
        rmap0 = ident
    elif initial == "affine":
        A = np.array([[1.001, 0.002, 0.8], [-0.0015, 0.999, -0.6]], f32)
        rmap0 = np.stack([A[0, 0] * ident[..., 0] + A[0, 1] * ident[..., 1] + A[0, 2],
                          A[1, 0] * ident[..., 0] + A[1, 1] * ident[..., 1] + A[1, 2]], -1).astype(f32)
    fo = oef.EccFlow(o)
    fo.set_reference_image(ref)
    want = fo.compute(cur, rmap0)
    fg = api.c_eccflow(_gpu_options(o))
    fg.set_reference_image(ref)
    got = fg.compute(cur, rmap0)
    _check("%dx%d %s" % (w, h, initial), got, want, fo, cur, rmap0)
    assert np.array_equal(fg.current_uv(), (got - ident).astype(f32)) or np.abs(fg.current_uv() - (got - ident)).max() <= 1e-4
    # the flow does what it is for: the refined map brings the frame onto the reference
    res_g = cv2.remap(cur, got, None, cv2.INTER_LINEAR)
    res_0 = cur if rmap0 is None else cv2.remap(cur, rmap0, None, cv2.INTER_LINEAR)
    assert np.abs(res_g - ref)[20:-20, 20:-20].mean() < 0.35 * np.abs(res_0 - ref)[20:-20, 20:-20].mean()


@pytest.mark.parametrize("method,kw", [(oef.DOWNSCALE_PYRAMID, dict(scale_factor=0.5)),
                                        (oef.DOWNSCALE_FULL_RESIZE, dict()),
                                        (oef.DOWNSCALE_RECURSIVE_RESIZE, dict(support_scale=3, max_iterations=2, update_multiplier=1.2)),
                                        (oef.DOWNSCALE_RECURSIVE_RESIZE, dict(max_pyramid_level=4, min_image_size=8)),
                                        (oef.DOWNSCALE_RECURSIVE_RESIZE, dict(scale_factor=0.5, support_scale=5, max_iterations=1, min_image_size=4))])
def test_eccflow_options_match_oracle(gpu, method, kw):
    from serstacker_b200 import api
    ref, cur, ident = _scene(240, 416, seed=11)
    o = oef.registration_options(downscale=method, **kw)
    fo = oef.EccFlow(o)
    fo.set_reference_image(ref)
    want = fo.compute(cur, ident)
    fg = api.c_eccflow(_gpu_options(o))
    fg.set_reference_image(ref)
    got = fg.compute(cur, ident)
    _check("method %d %s" % (method, kw), got, want, fo, cur, ident)


def test_eccflow_masks_match_oracle(gpu):
    """reference_mask (level masks by INTER_NEAREST) and input_mask (remapped with INTER_NEAREST / CONSTANT per iteration)."""
    from serstacker_b200 import api
    h, w = 270, 480
    ref, cur, ident = _scene(h, w, seed=5)
    rmask = np.full((h, w), 255, np.uint8)
    rmask[:, :37] = 0
    rmask[100:140, 200:260] = 0
    cmask = np.full((h, w), 255, np.uint8)
    cmask[-29:, :] = 0
    cmask[30:60, 300:380] = 0
    o = oef.registration_options()
    for rm, cm in [(rmask, None), (None, cmask), (rmask, cmask)]:
        fo = oef.EccFlow(o)
        fo.set_reference_image(ref, rm)
        want = fo.compute(cur, ident, cm)
        fg = api.c_eccflow(_gpu_options(o))
        fg.set_reference_image(ref, rm)
        got = fg.compute(cur, ident, cm)
        _check("masks ref=%s cur=%s" % (rm is not None, cm is not None), got, want, fo, cur, ident, cm)


def test_eccflow_full_size_1080p(gpu):
    """Config #2 geometry (1920 x 1080): 24 levels, 3 iterations each."""
    from serstacker_b200 import api
    ref, cur, ident = _scene(1080, 1920, seed=7, amp=60.0, smooth=40.0)
    o = oef.registration_options()
    fo = oef.EccFlow(o)
    fo.set_reference_image(ref)
    want = fo.compute(cur, ident)
    fg = api.c_eccflow(_gpu_options(o))
    fg.set_reference_image(ref)
    assert fg.num_levels() == len(fo.pyramid)
    got = fg.compute(cur, ident)
    _check("1920x1080", got, want, fo, cur, ident)


def test_eccflow_rejects_bad_arguments(gpu):
    from serstacker_b200 import api
    ref, cur, ident = _scene(64, 96, seed=9)
    f = api.c_eccflow(api.eccflow_options(registration_defaults=True))
    with pytest.raises(api.SskError):
        f.set_reference_image(np.zeros((64, 96, 3), f32))            # ecc2.cc:2498: single channel only
    f.set_reference_image(ref)
    with pytest.raises(api.SskError):
        f.compute(cur[:32], ident)                                   # size mismatch
    with pytest.raises(api.SskError):
        f.compute(cur, ident, np.zeros((10, 10), np.uint8))          # ecc2.cc:2695: mask size
    g = api.c_eccflow(api.eccflow_options(registration_defaults=True))
    with pytest.raises(api.SskError):
        g._shape = (64, 96)
        g.compute(cur, ident)                                        # ecc2.cc:2678: reference first


# ---------------------------------------------------------------------------------------------------------
# c_frame_registration with enable_eccflow_registration, and the stacking loop over per-pixel maps
# ---------------------------------------------------------------------------------------------------------
def _turbulent_sequence(w, h, n, seed, amp=30.0, smooth=18.0):
    """Planet frames (jittered, defocused) with a smooth per-frame turbulence warp on top."""
    from serstacker_b200 import synth
    frames, _, _ = synth.make_planet_sequence(w, h, n, seed=seed, radius=min(w, h) * 0.37, sigma_t=3.0, sigma_rot_deg=0.1,
                                              sigma_scale=0.001, blur_range=(0.8, 1.6), dtype="f32")
    rng = np.random.default_rng(seed + 100)
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    out = [frames[0]]
    for f in frames[1:]:
        du = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), smooth) * amp
        dv = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), smooth) * amp
        out.append(cv2.remap(f, xx + du, yy + dv, cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT101))
    return out


def _flow_registration_options(motion, method, interpolation):
    from oracle import registration as oreg
    o = oreg.ImageRegistrationOptions(motion_type=motion, interpolation=interpolation)
    o.ecc.ecc_method = method
    o.ecc.ecch_max_level = -1
    o.enable_eccflow_registration = True
    o.eccflow = oef.registration_options()
    return o


@pytest.mark.parametrize("motion", [0, 3])
def test_register_frame_with_eccflow_matches_oracle(gpu, motion):
    """c_frame_registration.cc:900-917: _current_remap = eccflow.compute(ecc_image, create_remap(transform), ecc_mask)."""
    from serstacker_b200 import api
    from oracle import registration as oreg
    frames = _turbulent_sequence(480, 270, 3, seed=21)
    oo = _flow_registration_options(motion, 3, cv2.INTER_LINEAR)
    ro = oreg.FrameRegistration(oo)
    ro.setup_reference_frame(frames[0])
    rg = api.c_frame_registration(api.registration_options(motion_type=motion, interpolation=1, enable_eccflow_registration=1,
                                                          ecc=dict(ecc_method=3, ecch_max_level=-1)))
    rg.setup_reference_frame(frames[0])
    rng = np.random.default_rng(5)
    for f in frames[1:]:
        assert ro.register_frame(f) and rg.register_frame(f)
        want, got = ro.current_remap, rg.current_remap()
        wimg, wmask = ro.remap(f, None)
        f2 = (f * (1 + rng.standard_normal(f.shape).astype(f32) * f32(1e-7))).astype(f32)
        assert ro.register_frame(f2)
        env = np.abs(ro.current_remap - want).max(axis=-1)
        d = np.abs(got - want).max(axis=-1)
        print("  register_frame + eccflow (motion %d): max |d map| = %.3g px, 99.9%% = %.3g, mean = %.3g  (oracle one-ulp envelope: max %.3g, 99.9%% %.3g, mean %.3g)"
              % (motion, d.max(), np.quantile(d, .999), d.mean(), env.max(), np.quantile(env, .999), env.mean()))
        assert np.quantile(d, .999) <= max(1e-3, 3 * np.quantile(env, .999))
        assert d.max() <= max(2e-3, 4 * env.max())
        # remap() through the refined map = the oracle's base_remap through its map (image where both masks agree)
        gimg, gmask = rg.remap(f)
        both = (wmask > 0) & (gmask > 0)
        assert (wmask != gmask).mean() < 1e-3
        gy_, gx_ = np.gradient(f)
        # cv::remap quantises the map to 1/32 px: a map difference can move a sample by one quantisation step
        assert np.abs(gimg - wimg)[both].max() <= (2 * d.max() + 1.0 / 32) * max(np.abs(gx_).max(), np.abs(gy_).max()) + 1e-5
        assert np.abs(gimg - wimg)[both].mean() <= 1e-4


@pytest.mark.parametrize("acc,interp", [(0, cv2.INTER_LINEAR), (1, cv2.INTER_CUBIC)])
def test_stack_with_eccflow_matches_oracle(gpu, acc, interp):
    """The per-frame loop with per-pixel maps (k_fused_flow): ECC -> eccflow -> warp through the flow map -> (weighted) average."""
    from serstacker_b200 import api
    from oracle import pipeline as opl
    frames = _turbulent_sequence(480, 270, 6, seed=31)
    so = opl.StackingOptions(accumulation_method=opl.ACC_WEIGHTED_AVERAGE if acc else opl.ACC_AVERAGE)
    so.registration = _flow_registration_options(3, 3, interp)
    avg_o, mask_o, _, _ = opl.run_stacking(frames, so)
    ro = api.registration_options(motion_type=3, interpolation=interp, enable_eccflow_registration=1, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=acc, max_batch=4))
    p.set_reference(frames[0])
    res = p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == len(frames)
    m = (mask_o > 0) & (mask_g > 0)
    rel = float(np.sqrt(((avg_g[m] - avg_o[m]) ** 2).sum()) / np.sqrt((avg_o[m] ** 2).sum()))
    print("  stack with eccflow (acc %d, interp %d): rel-L2 = %.3g, mask mismatch %.3g" % (acc, interp, rel, (mask_o != mask_g).mean()))
    assert rel <= 1e-4
    assert (mask_o != mask_g).mean() < 1e-3
    # and the flow did its job: the stack is sharper than the one registered without it
    so2 = opl.StackingOptions(accumulation_method=so.accumulation_method)
    so2.registration = _flow_registration_options(3, 3, interp)
    so2.registration.enable_eccflow_registration = False
    avg_n, _, _, _ = opl.run_stacking(frames, so2)
    lap = lambda im: float(np.abs(cv2.Laplacian(im, cv2.CV_32F))[40:-40, 40:-40].mean())
    assert lap(avg_g) > lap(avg_n)



def _bayer_turbulent_sequence(w, h, n, seed, amp=20.0, smooth=16.0):
    """Raw RGGB frames (uint16) of a textured colour scene: per frame a small shift plus a smooth turbulence warp."""
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.random((h + 40, w + 40)).astype(f32), (0, 0), 2.5)
    base = (base - base.min()) / (base.max() - base.min())
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    frames = []
    for i in range(n):
        tx, ty = (0.0, 0.0) if i == 0 else rng.normal(0.0, 1.5, 2)
        du = dv = 0
        if i:
            du = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), smooth) * amp
            dv = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), smooth) * amp
        lum = cv2.remap(base, xx + 20 + f32(tx) + du, yy + 20 + f32(ty) + dv, cv2.INTER_CUBIC)
        rgb = [0.15 + 0.70 * lum, 0.10 + 0.60 * lum, 0.20 + 0.45 * lum]
        mosaic = np.empty((h, w), f32)
        mosaic[0::2, 0::2] = rgb[0][0::2, 0::2]
        mosaic[0::2, 1::2] = rgb[1][0::2, 1::2]
        mosaic[1::2, 0::2] = rgb[1][1::2, 0::2]
        mosaic[1::2, 1::2] = rgb[2][1::2, 1::2]
        mosaic += rng.standard_normal(mosaic.shape).astype(f32) * f32(0.001)
        frames.append(np.clip(np.rint(mosaic * 65535.0), 0, 65535).astype(np.uint16))
    return frames, 16


def test_stack_bayer_average_with_eccflow_matches_oracle(gpu):
    """bayer_average with enable_eccflow_registration: the raw Bayer samples are gathered through the refined per-pixel
    map c_eccflow left in current_remap (c_frame_registration.cc:900-917, c_image_stacking_pipeline.cc:1750-1751) under the
    eroded mask of that map (k_fused_bayer with a flow field).  Two checks: (1) the gather itself is exact - the oracle's
    accumulator fed with the DEVICE maps reproduces the device stack bit for bit; (2) against the oracle's own maps the
    stack agrees to the eccflow envelope of the scene (printed)."""
    from serstacker_b200 import api
    from oracle import pipeline as opl, registration as oreg, accumulation as oacc
    from oracle.debayer import debayer_nn2
    frames, bpp = _bayer_turbulent_sequence(320, 224, 5, seed=41)
    oo = _flow_registration_options(0, 3, cv2.INTER_LINEAR)
    so = opl.StackingOptions(accumulation_method=opl.ACC_BAYER_AVERAGE)
    so.registration = oo
    avg_o, mask_o, _, _ = opl.run_bayer_stacking(frames, bpp, so, 8)
    rkw = dict(motion_type=0, interpolation=cv2.INTER_LINEAR, enable_eccflow_registration=1, ecc=dict(ecc_method=3, ecch_max_level=-1))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=api.registration_options(**rkw), accumulation_method=2,
                                                        bayer_colorid=8, max_batch=4))
    p.set_reference(frames[0], bpp=bpp)
    p.add_frames(frames)
    avg_g, mask_g = p.compute()
    assert p.accumulated_frames() == len(frames)
    # (1) device maps through the oracle's accumulator
    ro = oreg.FrameRegistration(oo)
    rg = api.c_frame_registration(api.registration_options(**rkw))
    bgr = [opl.to_float_frame(debayer_nn2(f, 8), bpp) for f in frames]
    ro.setup_reference_frame(bgr[0], None)
    rg.setup_reference_frame(bgr[0])
    acc = oacc.BayerAverage()
    acc.set_bayer_pattern(8)
    dmap = 0.0
    for f, raw in zip(bgr, frames):
        assert rg.register_frame(f) and ro.register_frame(f, None)
        mg = rg.current_remap()
        dmap = max(dmap, float(np.abs(mg - ro.current_remap).max()))
        _, mk = ro.custom_remap(mg, f, None, oo.interpolation, oo.border_mode, oo.border_value)
        acc.set_remap(mg)
        acc.add(opl.to_float_frame(raw, bpp), mk)
    avg_x, mask_x = acc.compute()
    assert np.array_equal(mask_g, mask_x) and np.array_equal(avg_g, avg_x)
    # (2) against the oracle's own maps
    m = (mask_o > 0) & (mask_g > 0)
    rel = float(np.sqrt(((avg_g[m] - avg_o[m]) ** 2).sum()) / np.sqrt((avg_o[m] ** 2).sum()))
    mism = float((mask_o != mask_g).mean())
    print("  bayer stack with eccflow: gather exact; max |d map| = %.3g px, rel-L2 vs the oracle's own maps = %.3g, mask mismatch %.3g"
          % (dmap, rel, mism))
    assert rel <= 1e-4
    assert mism < 1e-3
