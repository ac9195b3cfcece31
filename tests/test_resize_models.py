"""CPU: the cv::resize models of oracle/cvmodel.py (the arithmetic the c_eccflow and up-scaling kernels are written from) pinned
against cv2, and the CPU restatements of c_eccflow / c_canvas_average / the up-scaling helpers checked on known answers."""
import numpy as np
import cv2
import pytest

from oracle import cvmodel as m
from oracle import eccflow as oef
from oracle import pipeline as opl
from oracle.accumulation import CanvasAverage, WeightedAverage

f32 = np.float32


@pytest.mark.parametrize("ss,ds", [((17, 30), (270, 480)), ((3, 5), (6, 8)), ((270, 480), (3, 5)), ((203, 360), (270, 480)), ((1, 1), (5, 8))])
def test_resize_cubic_model_matches_cv2(ss, ds):
    rng = np.random.default_rng(ss[0] + ds[1])
    src = rng.standard_normal(ss + (2,)).astype(f32)
    want = cv2.resize(src, (ds[1], ds[0]), interpolation=cv2.INTER_CUBIC)
    got = m.resize_cubic_f32(src, (ds[1], ds[0]))
    assert np.abs(got - want).max() <= 4 * np.finfo(f32).eps * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("ss", [(270, 480), (203, 360), (153, 270), (5, 8), (810, 1440), (101, 77)])
def test_resize_area_table_model_matches_cv2(ss):
    rng = np.random.default_rng(ss[0])
    w, h = ss[1], ss[0]
    for _ in range(4):
        w, h = (w + 1) // 2, (h + 1) // 2
    src = rng.standard_normal(ss + (2,)).astype(f32)
    want = cv2.resize(src, (w, h), interpolation=cv2.INTER_AREA).reshape(h, w, 2)
    got = m.resize_area_tables_f32(src, (w, h))
    assert np.abs(got - want).max() <= 1.2e-7


@pytest.mark.parametrize("shape,k", [((37, 53), 1.5), ((64, 96), 1.5), ((37, 53), 3), ((135, 241), 1.5)])
def test_resize_linear_models_match_cv2_without_ipp(shape, k):
    rng = np.random.default_rng(shape[1])
    src = rng.standard_normal(shape).astype(f32)
    h, w = shape
    dsize = (w * 3 // 2, h * 3 // 2) if k == 1.5 else (w * 3, h * 3)
    want = opl._resize_no_ipp(src, dsize, cv2.INTER_LINEAR if k == 1.5 else cv2.INTER_LINEAR_EXACT)
    got = m.resize_linear_f32(src, dsize)
    assert np.abs(got - want).max() <= 1.2e-7 * max(1.0, float(np.abs(want).max())) and (got == want).mean() > 0.99
    if k == 1.5:
        mask = ((rng.random(shape) > 0.1) * 255).astype(np.uint8)
        wantm = opl._resize_no_ipp(mask, dsize, cv2.INTER_LINEAR) >= 255
        assert np.array_equal(m.resize_linear_u8_ge255(mask, dsize), wantm)


def _turbulent_pair(h, w, seed):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.random((h + 40, w + 40)).astype(f32), (0, 0), 2.0)
    base = (base - base.min()) / (base.max() - base.min())
    ref = base[20:20 + h, 20:20 + w].copy()
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    du = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), 25) * 40
    dv = cv2.GaussianBlur(rng.standard_normal((h, w)).astype(f32), (0, 0), 25) * 40
    cur = cv2.remap(base, xx + 20 + du, yy + 20 + dv, cv2.INTER_CUBIC)
    return ref, cur, np.stack([xx, yy], -1), np.stack([du, dv], -1)


def test_eccflow_oracle_recovers_a_smooth_warp():
    """Known answer: a frame warped by a smooth field d is mapped back by rmap ~ identity + d (up to the aperture of the 16-px
    support), and the residual against the reference drops by more than half an order of magnitude."""
    ref, cur, ident, d = _turbulent_pair(180, 260, 3)
    f = oef.EccFlow(oef.registration_options())
    f.set_reference_image(ref)
    sizes = [e.reference_image.shape for e in f.pyramid]
    assert sizes[0] == (180, 260) and sizes[1] == (135, 195) and max(sizes[-1]) <= 6      # (int)((w + 1) * 0.75) recursion
    rmap = f.compute(cur, ident)
    res = cv2.remap(cur, rmap, None, cv2.INTER_LINEAR)
    c = (slice(24, -24), slice(24, -24))
    assert np.abs(res - ref)[c].mean() < 0.25 * np.abs(cur - ref)[c].mean()
    # ref(x) = base(x); cur(x) = base(x + d(x))  =>  cur(x - d) ~ ref(x): the flow is ~ -d
    assert np.abs(f.uv + d)[c].mean() < 0.3 * np.abs(d)[c].mean()
    # an empty initial map and the identity map are the same start (ecc2.cc:2783-2800)
    assert np.array_equal(f.compute(cur, None), rmap)


def test_eccflow_oracle_pyramid_rules():
    o = oef.EccFlowOptions(downscale=oef.DOWNSCALE_PYRAMID, scale_factor=0.5, min_image_size=4)
    sizes = [s for s, _ in oef.EccFlow(o).level_sizes((480, 270))]
    assert sizes[:3] == [(480, 270), (240, 135), (120, 68)] and min(sizes[-1]) > 4
    # big aspect ratio: the small levels are reduced from level 0 (ecc2.cc:2553-2561)
    lv = oef.EccFlow(oef.registration_options()).level_sizes((600, 40))
    assert all(src == 0 for s, src in lv[1:] if min(s) <= 5) and all(src == i for i, (s, src) in enumerate(lv[1:]) if min(s) > 5)
    assert oef.EccFlow(oef.registration_options(max_pyramid_level=2)).level_sizes((480, 270))[-1][0] == (270, 153)


def test_canvas_average_oracle_equals_weighted_average_inside_the_box():
    """Without a remap c_canvas_average is c_weigthed_average on the box of the first frame; a shift of the box by an integer
    offset with the identity map places the frame there."""
    rng = np.random.default_rng(1)
    frames = [rng.random((40, 60)).astype(f32) for _ in range(4)]
    c, w = CanvasAverage(), WeightedAverage()
    for fr in frames:
        assert c.add(fr) and w.add(fr)
    x, y, bw, bh = c.last_bbox
    assert (bw, bh) == (60, 40) and c.accumulator.shape[:2] == (60, 90)
    assert np.array_equal(c.compute((x, y, bw, bh))[0], w.compute()[0])
    yy, xx = np.mgrid[0:40, 0:60].astype(f32)
    c2 = CanvasAverage(canvas_size=(300, 200))
    c2.add(frames[0])
    x, y = c2.last_bbox[:2]
    c2.add(frames[1], None, np.stack([xx, yy], -1), (x + 50, y + 30, 60, 40))
    avg, mask = c2.compute()
    assert np.array_equal(avg[y + 30:y + 70, x + 60:x + 110], frames[1][:, 10:])       # beyond the first frame's box
    assert mask[y + 35, x + 100] == 255 and mask[0, 0] == 0
    with pytest.raises(ValueError):
        small = CanvasAverage(canvas_size=(100, 200))          # the 64-px shift pushes the box out of a canvas this narrow
        small.add(frames[0])
        small.add(frames[1], None, np.stack([xx, yy], -1), (5, small.last_bbox[1], 60, 40))


def test_upscale_oracle_shapes_and_mask_rule():
    rng = np.random.default_rng(2)
    img = rng.random((20, 30)).astype(f32)
    mask = np.full((20, 30), 255, np.uint8)
    mask[8:12, 10:14] = 0
    for opt, shape in [(opl.UPSCALE_PYRUP, (40, 60)), (opl.UPSCALE_X15, (30, 45)), (opl.UPSCALE_X30, (60, 90))]:
        d, dm = opl.upscale_image(opt, img, mask)
        assert d.shape == shape and dm.shape == shape and set(np.unique(dm)) <= {0, 255}
        assert dm[0, 0] == 255 and dm[shape[0] // 2, int(12 * shape[1] / 30)] == 0
    yy, xx = np.mgrid[0:20, 0:30].astype(f32)
    flow = np.stack([np.full((20, 30), 2, f32), np.full((20, 30), -1, f32)], -1)
    assert np.allclose(opl.upscale_optflow(opl.UPSCALE_X15, flow), [3.0, -1.5])
    assert np.array_equal(opl.upscale_remap(opl.UPSCALE_NONE, flow), flow)
