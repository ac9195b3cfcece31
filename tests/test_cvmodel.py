"""CPU: the numpy model of cv::remap (oracle/cvmodel.py) - the specification the CUDA samplers are written
to - against cv2 itself."""
import numpy as np
import cv2
import pytest

from oracle import cvmodel as m
from oracle import transforms as tf


def _maps(rng, w, h, n=4):
    t = tf.AffineTransform()
    for _ in range(n):
        p = np.array([1 + rng.normal(0, 0.02), rng.normal(0, 0.02), rng.normal(0, 5),
                      rng.normal(0, 0.02), 1 + rng.normal(0, 0.02), rng.normal(0, 5)], np.float32)
        yield t.create_remap((w, h), p)


@pytest.mark.parametrize("interp,ci", [("linear", cv2.INTER_LINEAR), ("cubic", cv2.INTER_CUBIC), ("nearest", cv2.INTER_NEAREST)])
@pytest.mark.parametrize("border", [m.BORDER_REPLICATE, m.BORDER_REFLECT101, m.BORDER_CONSTANT, m.BORDER_REFLECT])
def test_remap_f32_model_matches_cv2(interp, ci, border):
    rng = np.random.default_rng(1)
    h, w = 47, 61
    src = rng.random((h, w)).astype(np.float32)
    for rmap in _maps(rng, w, h):
        ref = cv2.remap(src, rmap, None, ci, borderMode=border, borderValue=0.25)
        mine = m.remap_f32(src, rmap, interp, border, 0.25)
        if interp == "cubic":
            assert np.abs(ref - mine).max() <= 1e-6
        else:
            assert np.array_equal(ref, mine)      # bilinear / nearest: bit-exact


@pytest.mark.parametrize("interp,ci", [("linear", cv2.INTER_LINEAR), ("nearest", cv2.INTER_NEAREST)])
def test_remap_transparent_model_matches_cv2(interp, ci):
    rng = np.random.default_rng(2)
    h, w = 40, 52
    src = rng.random((h, w)).astype(np.float32)
    for rmap in _maps(rng, w, h):
        dst0 = np.full((h, w), 7.0, np.float32)
        ref = cv2.remap(src, rmap, None, ci, dst=dst0.copy(), borderMode=cv2.BORDER_TRANSPARENT)
        mine = m.remap_f32(src, rmap, interp, m.BORDER_TRANSPARENT, 0.0, dst=dst0)
        assert np.abs(ref - mine).max() <= 2e-7


@pytest.mark.parametrize("interp,ci", [("linear", cv2.INTER_LINEAR), ("cubic", cv2.INTER_CUBIC)])
def test_mask_fixed_point_model_matches_cv2(interp, ci):
    rng = np.random.default_rng(3)
    h, w = 45, 57
    for rmap in _maps(rng, w, h, n=6):
        ref = cv2.remap(np.full((h, w), 255, np.uint8), rmap, None, ci, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        for th in (255, 254, 250):
            v, val = m.remap_u8_all255_valid((w, h), rmap, interp, th)
            assert np.array_equal(ref >= th, v)
            assert np.array_equal(ref.astype(int), val)


def test_bilinear_all255_thresholds_coincide():
    """With an all-255 source the >=250 / >=254 / >=255 tests select the same pixels (bilinear): the CUDA
    kernels use one closed form for all three call sites (ecc2.cc:128, 215, 1312)."""
    rng = np.random.default_rng(4)
    h, w = 33, 41
    for rmap in _maps(rng, w, h, n=8):
        a, _ = m.remap_u8_all255_valid((w, h), rmap, "linear", 255)
        b, _ = m.remap_u8_all255_valid((w, h), rmap, "linear", 254)
        c, _ = m.remap_u8_all255_valid((w, h), rmap, "linear", 250)
        assert np.array_equal(a, b) and np.array_equal(a, c)
        ix, fx = m.quantize(rmap[..., 0])
        iy, fy = m.quantize(rmap[..., 1])
        closed = (ix >= 0) & (iy >= 0) & (ix < w) & (iy < h) & ((fx == 0) | (ix + 1 < w)) & ((fy == 0) | (iy + 1 < h))
        assert np.array_equal(a, closed)


def test_small_matrix_models_bit_exact():
    """The device-side solver algebra is written to these models: hal::Cholesky32f, cv::invertAffineTransform."""
    rng = np.random.default_rng(5)
    for M in (4, 6, 8):
        for _ in range(50):
            J = rng.normal(size=(50, M)) * np.array([300, 200, 1, 300, 200, 1, 90000, 60000][:M])
            H = (J.T @ J).astype(np.float32)
            v = (J.T @ rng.normal(size=(50, 1))).astype(np.float32)
            ok, x = cv2.solve(H, v, flags=cv2.DECOMP_CHOLESKY)
            ok2, x2 = m.chol_solve_f32(H, v)
            assert bool(ok) == ok2 and np.array_equal(x, x2)
            oki, Hi = cv2.invert(H, flags=cv2.DECOMP_CHOLESKY)
            ok3, Hi2 = m.chol_solve_f32(H, np.eye(M, dtype=np.float32))
            assert (oki != 0) == ok3 and np.array_equal(Hi, Hi2)
    for _ in range(500):
        s = rng.choice([0.01, 0.2, 1.0])
        A = np.array([[1 + rng.normal(0, s), rng.normal(0, s), rng.normal(0, 50)],
                      [rng.normal(0, s), 1 + rng.normal(0, s), rng.normal(0, 50)]], np.float32)
        assert np.array_equal(cv2.invertAffineTransform(A), m.invert_affine_f32(A))


def test_pyrdown_model_bit_exact():
    rng = np.random.default_rng(6)
    for (h, w, ds) in [(64, 80, None), (61, 83, None), (240, 320, None), (270, 480, (240, 134)), (33, 47, None), (134, 240, (120, 66))]:
        src = rng.random((h, w)).astype(np.float32)
        ref = cv2.pyrDown(src) if ds is None else cv2.pyrDown(src, dstsize=ds)
        assert np.array_equal(ref, m.pyrdown_f32(src, ds)), (h, w, ds)


def test_pyrup_model():
    """cv::pyrUp restated (ecc_normalize / lpg up-scaling): bit-exact for even, odd and +1 destination sizes."""
    rng = np.random.default_rng(9)
    for (h, w, dh, dw) in [(30, 40, 60, 80), (31, 41, 61, 81), (31, 41, 62, 82), (17, 23, 33, 45), (64, 96, 128, 192), (5, 4, 9, 8)]:
        a = rng.random((h, w)).astype(np.float32)
        assert np.array_equal(m.pyrup_f32(a, (dw, dh)), cv2.pyrUp(a, dstsize=(dw, dh))), (h, w, dh, dw)


def test_pyrdown_replicate_model_is_border_independent_inside():
    """ecc_downscale uses BORDER_REPLICATE: interior outputs equal the REFLECT101 ones; only the border taps differ."""
    rng = np.random.default_rng(10)
    a = rng.random((41, 57)).astype(np.float32)
    r = cv2.pyrDown(a, borderType=cv2.BORDER_REPLICATE)
    d = cv2.pyrDown(a)
    assert np.array_equal(r[1:-1, 1:-1], d[1:-1, 1:-1])


@pytest.mark.parametrize("sw,sh,dw,dh", [(97, 61, 49, 31), (100, 60, 75, 45), (101, 77, 25, 19), (50, 40, 37, 29), (96, 64, 24, 16), (99, 63, 33, 21)])
def test_resize_area_model_bit_exact(sw, sh, dw, dh):
    """INTER_AREA down-scaling, fractional ratios (ResizeArea) and integer ratios 3 / 4 (ResizeAreaFast): bit-exact."""
    rng = np.random.default_rng(sw * 1000 + dw)
    src = rng.random((sh, sw)).astype(np.float32)
    want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
    assert np.array_equal(m.resize_area_f32(src, (dw, dh)), want)


@pytest.mark.parametrize("sw,sh,f", [(402, 302, 0.25), (80, 60, 0.75), (101, 77, 0.4)])
def test_resize_area_model_with_fx_fy(sw, sh, f):
    """scaleImage's call form: dsize derived from fx = fy (cells clipped by the edge when cvRound rounds up)."""
    rng = np.random.default_rng(sw)
    src = rng.random((sh, sw)).astype(np.float32)
    want = cv2.resize(src, (0, 0), fx=f, fy=f, interpolation=cv2.INTER_AREA)
    got = m.resize_area_f32(src, (want.shape[1], want.shape[0]), inv_scale=(f, f))
    assert np.array_equal(got, want)


def test_resize_area_model_2x2_simd_form():
    """2 x 2 cells: the SIMD columns use (s00 + s01) + (s10 + s11), the tail columns the scalar running form; which
    columns are 'tail' depends on the vector width of the OpenCV build, so one of the common widths must match."""
    rng = np.random.default_rng(5)
    src = rng.random((50, 70)).astype(np.float32)
    want = cv2.resize(src, (35, 25), interpolation=cv2.INTER_AREA)
    assert any(np.array_equal(m.resize_area_f32(src, (35, 25), simd_lanes=l), want) for l in (4, 8, 16))


def test_remap_linear_u8_fixed_point_model_is_bit_exact():
    """The 8-bit bilinear mask remap the ECC solvers threshold at 255 / 254 / 250 (binary and non-binary mask values)."""
    import cv2
    from oracle import cvmodel
    rng = np.random.default_rng(0)
    h, w = 40, 56
    m = (rng.random((h, w)) > 0.3).astype(np.uint8) * 255
    m[5:9, 10:20] = 128
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for tx, ty in [(0.3, -0.7), (2.53, 1.01), (-3.2, 4.96), (0.03125, 0.03125), (0.0, 0.0)]:
        mapx = (xx * 1.01 + tx).astype(np.float32)
        mapy = (yy * 0.99 + ty + 0.002 * xx).astype(np.float32)
        want = cv2.remap(m, np.dstack([mapx, mapy]), None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(cvmodel.remap_linear_u8(m, mapx, mapy), want)


def test_lanczos4_models_match_cv2():
    """cv::remap(INTER_LANCZOS4): the float sampler (footprint inside the image) and the fixed-point all-255 validity rule."""
    rng = np.random.default_rng(12)
    src = rng.random((60, 70)).astype(np.float32)
    yy, xx = np.mgrid[0:60, 0:70].astype(np.float32)
    for k in range(3):
        mx = (xx * np.float32(1.01 - 0.01 * k) - np.float32(3.3 + k) + yy * np.float32(0.02)).astype(np.float32)
        my = (yy * np.float32(0.99) + np.float32(2.7 - 2 * k) - xx * np.float32(0.01)).astype(np.float32)
        want = cv2.remap(src, mx, my, cv2.INTER_LANCZOS4)
        got = m.remap_lanczos4_f32(src, mx, my)
        ok = ~np.isnan(got)
        assert ok.sum() > 1000 and np.array_equal(got[ok], want[ok])
        msk = np.full((60, 70), 255, np.uint8)
        wv = cv2.remap(msk, mx, my, cv2.INTER_LANCZOS4, borderMode=cv2.BORDER_CONSTANT, borderValue=0) >= 255
        assert np.array_equal(m.remap_lanczos4_all255_valid((60, 70), mx, my), wv)
