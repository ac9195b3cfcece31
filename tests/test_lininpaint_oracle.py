"""CPU: the vectorised oracle of linear_interpolation_inpaint against a loop-by-loop transcription of the reference's row
pass (_interpolate_holes_h2, linear_interpolation_inpaint.cc:14-117), and the properties the fill must have."""
import numpy as np

from oracle import inpaint as oi

f32 = np.float32


def _rows_literal(img, mask):
    h, w = mask.shape
    out, dist = img.copy(), np.zeros((h, w), f32)
    for y in range(h):
        start = 0
        while True:
            while start < w and mask[y, start]:
                start += 1
            if start >= w:
                break
            end = start + 1
            while end < w and not mask[y, end]:
                end += 1
            if start > 0 and end < w:
                s, e = start - 1, end
                scale = f32(1.0) / f32(end - start)
                sv, ev = img[y, s], img[y, e]
                kk = ((ev - sv) * scale).astype(f32)
                for x in range(start, e):
                    out[y, x] = (sv + f32(x - s) * kk).astype(f32)
                    dist[y, x] = max(x - s, e - x)
            elif start > 0:
                s = start - 1
                for x in range(start, w):
                    out[y, x] = img[y, s]
                    dist[y, x] = x - s
            elif end < w:
                for x in range(start, end):
                    out[y, x] = img[y, end]
                    dist[y, x] = end - x
            start = end
            if start >= w:
                break
    return out, dist


def test_row_pass_equals_literal_transcription():
    rng = np.random.default_rng(0)
    img = rng.random((20, 33, 2)).astype(f32)
    mask = (rng.random((20, 33)) > 0.4).astype(np.uint8) * 255
    mask[3] = 0
    mask[5, :10] = 0
    mask[7, 20:] = 0
    a, da = oi._interpolate_holes_1d(img, mask)
    b, db = _rows_literal(img, mask)
    assert np.array_equal(a, b) and np.array_equal(da, db)


def test_fill_properties():
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:40, 0:56].astype(f32)
    ramp = (0.01 * xx + 0.02 * yy).astype(f32)
    mask = (rng.random(ramp.shape) > 0.3).astype(np.uint8) * 255
    mask[0, 0] = 255
    out = oi.linear_interpolation_inpaint(ramp, mask)
    assert np.array_equal(out[mask > 0], ramp[mask > 0])            # valid pixels untouched
    inner = np.zeros_like(mask, bool)
    inner[5:-5, 5:-5] = True
    assert np.isfinite(out).all()
    assert np.abs(out - ramp)[inner].max() < 0.05                   # a linear ramp is reproduced up to the reference's one-sided runs
    assert np.array_equal(oi.linear_interpolation_inpaint(ramp, None), ramp)
    empty = np.zeros_like(mask)
    assert np.array_equal(oi.linear_interpolation_inpaint(ramp, empty), ramp)   # nothing to interpolate from: unchanged
