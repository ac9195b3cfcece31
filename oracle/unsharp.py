"""
Oracle restatement of unsharp_mask (core/proc/unsharp_mask.cc:14-118), applied by the stacking pipeline to the
master / reference frame before registration starts (c_image_stacking_pipeline.cc:1302-1306; defaults
unsharp_sigma = 1, unsharp_alpha = 0.8: c_image_stacking_pipeline.h:169-170).

  create_lpass_image   unsharp_mask.cc:14-69   (sigma <= 2: sepFilter2D with an 2*max(1,int(5 sigma))+1 tap Gaussian,
                                                BORDER_REFLECT; larger sigma: pyrDown chain + residual blur + pyrUp chain)
  unsharp_mask         unsharp_mask.cc:72-118  (addWeighted(src, 1/(1-alpha), lpass, -alpha/(1-alpha)), optional clamp)

Every pixel primitive is delegated to cv2.  Test infrastructure only (see oracle/__init__.py).
"""
import math
import numpy as np
import cv2

f32 = np.float32
_BORDER = cv2.BORDER_REFLECT


def _gaussian_blur(src, sigma):
    # unsharp_mask.cc:19-23
    G = cv2.getGaussianKernel(2 * max(1, int(sigma * 5)) + 1, sigma, cv2.CV_32F)
    return cv2.sepFilter2D(src, -1, G, G, borderType=_BORDER)


def lpass_pyramid_level(rows, cols, sigma):
    # unsharp_mask.cc:25-42 -> (pyramid_level, Ci)
    level, Ci = 0, 0
    if sigma > 2:
        m, imax = min(rows, cols), 0
        while m >> 1:
            m >>= 1
            imax += 1
        C = int(sigma * sigma / 2)
        while level < imax and (1 + 4 * Ci) <= C:
            Ci = 1 + 4 * Ci
            level += 1
    return level, Ci


def create_lpass_image(src, sigma):
    level, Ci = lpass_pyramid_level(src.shape[0], src.shape[1], sigma)
    if level < 1:
        return _gaussian_blur(src, sigma)
    delta = math.sqrt(sigma * sigma - 2 * Ci) / (1 << level)
    sizes = [(src.shape[1], src.shape[0])]
    lp = cv2.pyrDown(src, borderType=_BORDER)
    for _ in range(1, level):
        sizes.append((lp.shape[1], lp.shape[0]))
        lp = cv2.pyrDown(lp, borderType=_BORDER)
    if delta > 0:
        lp = _gaussian_blur(lp, delta)
    for sz in reversed(sizes):
        lp = cv2.pyrUp(lp, dstsize=sz)
    return lp


def unsharp_mask(src, sigma, alpha, outmin=-1.0, outmax=-1.0):
    """unsharp_mask(src, dst, sigma, alpha, outmin, outmax) for CV_32F images (no implicit clamp for float depth)."""
    src = np.ascontiguousarray(src, dtype=f32)
    if sigma <= 0 or alpha <= 0:
        dst = src.copy()
    else:
        lp = create_lpass_image(src, sigma)
        dst = cv2.addWeighted(src, 1.0 / (1.0 - alpha), lp, -alpha / (1.0 - alpha), 0)
    if outmax > outmin:
        dst = np.maximum(np.minimum(dst, f32(outmax)), f32(outmin)).astype(f32)   # cv::min / cv::max with a scalar
    return dst
