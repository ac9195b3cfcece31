"""
Oracle restatement of median_filter_bad_pixels (non-Bayer branch), /root/reference/core/proc/bad_pixels.cc:14-70:
cv::medianBlur(5), cv::absdiff, cv::boxFilter(5 x 5, normalised, BORDER_DEFAULT), then the thresholded replacement;
and of bayer_denoise (returnBayerPlanes = false), /root/reference/core/io/debayer.cc:1471-1611: the same test with 3 x 3
windows on the four colour planes of the raw mosaic (_extract_bayer_planes, debayer.cc:96-135), written back in place.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
import cv2

f32 = np.float32


def median_filter_bad_pixels(image, variation_threshold):
    """-> filtered copy.  image: uint8 / uint16 / float32, HxW or HxWxC (C = 3 or 4 as cv::medianBlur accepts)."""
    median = cv2.medianBlur(image, 5)
    mad = cv2.absdiff(image, median)
    mad = cv2.boxFilter(mad, -1, (5, 5), anchor=(2, 2), normalize=True, borderType=cv2.BORDER_DEFAULT)
    mv = f32(1.0) if image.dtype != np.float32 else f32(1.0 / 256.0)
    k = f32(variation_threshold)
    p, m = image.astype(f32), median.astype(f32)
    bad = np.abs(m - p) > (k * mad.astype(f32)).astype(f32) + mv
    out = image.copy()
    out[bad] = median[bad]
    return out


def bayer_denoise(raw, variation_threshold):
    """-> filtered copy of a raw single-channel Bayer mosaic (even size).  debayer.cc:1507-1590."""
    assert raw.ndim == 2 and raw.shape[0] % 2 == 0 and raw.shape[1] % 2 == 0
    planes = np.ascontiguousarray(np.stack([raw[0::2, 0::2], raw[0::2, 1::2], raw[1::2, 0::2], raw[1::2, 1::2]], axis=-1))
    median = cv2.medianBlur(planes, 3)
    mad = cv2.absdiff(planes, median)
    mad = cv2.boxFilter(mad, -1, (3, 3), anchor=(-1, -1), normalize=True, borderType=cv2.BORDER_DEFAULT)
    minvar = f32(1.0) if raw.dtype != np.float32 else f32(1.0 / 256.0)
    k = f32(variation_threshold)
    p, m = planes.astype(f32), median.astype(f32)
    bad = np.abs(m - p) > (k * mad.astype(f32)).astype(f32) + minvar
    out = raw.copy()
    for c, (oy, ox) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        view = out[oy::2, ox::2]
        view[bad[..., c]] = median[..., c][bad[..., c]]
    return out
