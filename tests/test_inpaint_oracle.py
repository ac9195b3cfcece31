"""CPU checks of the average_pyramid_inpaint oracle (oracle/inpaint.py): its box-sum model against cv2, and the fused,
always-to-the-bottom form the CUDA kernels implement (serstacker_b200/csrc/ssk_inpaint.cu) against the reference's
recursive form (core/proc/inpaint/average_pyramid_inpaint.cc:69-95)."""
import numpy as np
import cv2
import pytest

from oracle import inpaint as oip

f32 = np.float32


def holes_mask(rng, rows, cols, fill=0.75, boxes=2):
    m = (rng.random((rows, cols)) < fill).astype(np.uint8) * 255
    for _ in range(boxes):
        h, w = rng.integers(2, max(3, rows // 3)), rng.integers(2, max(3, cols // 3))
        y, x = rng.integers(0, rows - h + 1), rng.integers(0, cols - w + 1)
        m[y:y + h, x:x + w] = 0
    return m


def _clamped_window_sum(img, msk):
    s = oip.box_sum_model(img)
    c = oip.box_sum_model(msk)
    return s, c


def kernel_model(src, mask, max_levels=100):
    """The device algorithm: level 0 = masked copy; down pass evaluates the filtered level only on the kept
    (odd, clamped) pixels; every level is descended (a full mask makes the deeper levels no-ops); the up pass gathers
    the odd/odd pixels of the 3x3 window from the level below."""
    img = np.where((mask != 0)[(...,) + (None,) * (src.ndim - 2)], src, f32(0)).astype(f32)
    msk = (mask.astype(f32) * f32(1.0 / 255.0)).astype(f32)
    levels = [(img, msk)]
    while min(levels[-1][0].shape[:2]) > 1 and len(levels) <= max_levels:
        I, M = levels[-1]
        s, c = _clamped_window_sum(I, M)
        fb = M != 0
        have = ~fb & (c != 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            sc = (f32(1) / c).astype(f32)
        out = s.copy()
        scb = sc[(...,) + (None,) * (src.ndim - 2)]
        with np.errstate(invalid="ignore"):
            out[have] = (s * scb).astype(f32)[have]
        out[fb] = I[fb]
        om = np.where(fb | have, f32(1), c).astype(f32)
        levels.append((oip.downstrike_even(out), oip.downstrike_even(om)))
    L = levels[-1][0]
    for I, M in reversed(levels[:-1]):
        up, z = oip.upject_even(L, (I.shape[1], I.shape[0]))
        s, c = _clamped_window_sum(up, z)
        fb = M != 0
        have = ~fb & (c != 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            sc = (f32(1) / c).astype(f32)
        out = s.copy()
        scb = sc[(...,) + (None,) * (src.ndim - 2)]
        with np.errstate(invalid="ignore"):
            out[have] = (s * scb).astype(f32)[have]
        out[fb] = I[fb]
        Mout = np.where(fb | have, f32(1), c).astype(f32)
        L = out
    return L, np.clip(np.rint(Mout * f32(255)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("shape", [(37, 53), (64, 64), (101, 77, 3), (135, 240)])
def test_box_sum_model_matches_cv2(shape):
    rng = np.random.default_rng(1)
    a = rng.random(shape, dtype=f32)
    a[rng.random(shape[:2]) < 0.3] = 0
    a *= (10.0 ** rng.integers(-6, 1, size=shape)).astype(f32)
    ref = cv2.boxFilter(a, -1, (3, 3), normalize=False, borderType=cv2.BORDER_REPLICATE)
    assert np.array_equal(ref, oip.box_sum_model(a))


@pytest.mark.parametrize("shape", [(2, 2), (5, 7), (33, 47), (64, 96), (135, 240), (90, 61, 3)])
@pytest.mark.parametrize("fill", [0.02, 0.5, 0.9])
def test_fused_form_equals_recursive_form(shape, fill):
    rng = np.random.default_rng(shape[0] * 131 + int(fill * 100))
    src = rng.random(shape, dtype=f32)
    mask = holes_mask(rng, shape[0], shape[1], fill)
    if mask.all():
        mask[0, 0] = 0
    a, am = oip.average_pyramid_inpaint(src, mask)
    b, bm = kernel_model(src, mask)
    assert np.array_equal(a, b)
    assert np.array_equal(am, bm)


def test_max_levels_and_valid_pixels_kept():
    rng = np.random.default_rng(5)
    src = rng.random((80, 120), dtype=f32)
    mask = holes_mask(rng, 80, 120, 0.6)
    for ml in (1, 2, 3):
        a, am = oip.average_pyramid_inpaint(src, mask, ml)
        b, bm = kernel_model(src, mask, ml)
        assert np.array_equal(a, b) and np.array_equal(am, bm)
        assert np.array_equal(a[mask > 0], src[mask > 0])
    full, fm = oip.average_pyramid_inpaint(src, mask)
    assert fm.min() == 255


def test_full_and_empty_masks():
    rng = np.random.default_rng(6)
    src = rng.random((31, 45), dtype=f32)
    full = np.full((31, 45), 255, np.uint8)
    a, am = oip.average_pyramid_inpaint(src, full)
    assert np.array_equal(a, src) and np.array_equal(am, full)
    none = np.zeros((31, 45), np.uint8)
    a, am = oip.average_pyramid_inpaint(src, none)
    assert not a.any() and am.min() == 255     # the 1-px bottom level is up-jected as valid zeros
    b, bm = kernel_model(src, none)
    assert np.array_equal(a, b) and np.array_equal(am, bm)
