"""GPU parity: median_filter_bad_pixels (core/proc/bad_pixels.cc:14-70, non-Bayer branch) against oracle/badpixels.py
(cv2.medianBlur / absdiff / boxFilter).  Integer depths are exact; for CV_32F the 5 x 5 mean of |image - median| is a double
sum here and a running double sum in cv::boxFilter, so a replacement decision can differ only on an exact tie of the
threshold: the test allows 1e-5 of the samples."""
import numpy as np
import cv2
import pytest

from oracle import badpixels as obp

pytestmark = pytest.mark.gpu


def _frame(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.random(shape).astype(np.float32), (0, 0), 2.0)
    base = base.reshape(shape)
    img = base + rng.standard_normal(shape).astype(np.float32) * 0.01
    hot = rng.random(shape) < 0.002
    img[hot] += rng.uniform(0.3, 0.6, int(hot.sum())).astype(np.float32)      # hot pixels
    cold = rng.random(shape) < 0.001
    img[cold] = 0.0
    img = np.clip(img, 0, 1)
    if dtype == np.float32:
        return img.astype(np.float32), int(hot.sum())
    mx = np.iinfo(dtype).max
    return np.rint(img * mx).astype(dtype), int(hot.sum())


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("shape", [(97, 131), (64, 80, 3), (270, 480)])
@pytest.mark.parametrize("k", [3.0, 6.0])
def test_median_filter_bad_pixels_matches_oracle(gpu, dtype, shape, k):
    from serstacker_b200 import api
    img, nhot = _frame(shape, dtype, seed=shape[1] + int(k))
    want = obp.median_filter_bad_pixels(img, k)
    got = api.median_filter_bad_pixels(img, k)
    assert got.dtype == img.dtype and got.shape == img.shape
    ndiff = int((got != want).sum())
    assert int((want != img).sum()) >= nhot // 2                      # the filter did replace the planted outliers
    if dtype == np.float32:
        assert ndiff <= max(1, int(1e-5 * img.size)), ndiff
        same = got == want
        assert np.array_equal(got[same], want[same])
    else:
        assert ndiff == 0


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("shape", [(96, 132), (300, 480)])
@pytest.mark.parametrize("k", [3.0, 6.0])
def test_bayer_denoise_matches_oracle(gpu, dtype, shape, k):
    from serstacker_b200 import api
    img, nhot = _frame(shape, dtype, seed=shape[0] + int(k))
    # a mosaic: the four colour planes at different gains, as a CFA delivers them
    gain = np.array([[1.0, 0.6], [0.6, 0.35]], np.float32)
    mos = img.astype(np.float32) * np.tile(gain, (shape[0] // 2, shape[1] // 2))
    mos = mos.astype(np.float32) if dtype == np.float32 else np.rint(mos).astype(dtype)
    want = obp.bayer_denoise(mos, k)
    got = api.bayer_denoise(mos, k)
    assert got.dtype == mos.dtype and got.shape == mos.shape
    assert int((want != mos).sum()) >= nhot // 4
    ndiff = int((got != want).sum())
    if dtype == np.float32:
        assert ndiff <= max(1, int(1e-5 * mos.size)), ndiff
    else:
        assert ndiff == 0


def test_bayer_denoise_rejects_uneven(gpu):
    from serstacker_b200 import api
    with pytest.raises(api.SskError):
        api.bayer_denoise(np.zeros((7, 8), np.uint16), 3.0)


def test_median_filter_bad_pixels_rejects(gpu):
    from serstacker_b200 import api
    with pytest.raises(api.SskError):
        api.median_filter_bad_pixels(np.zeros((1, 5), np.float32), 3.0)
