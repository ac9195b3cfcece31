"""
Oracle restatement of the finishing step of the stacking pass: average_pyramid_inpaint.

  average_pyramid_inpaint      core/proc/inpaint/average_pyramid_inpaint.cc:97-127
  average_pyramid_recurse      core/proc/inpaint/average_pyramid_inpaint.cc:69-95
  _average_pyramid_filter2     core/proc/inpaint/average_pyramid_inpaint.cc:17-56
  downstrike_even              core/proc/downstrike.cc:34-78
  upject_even                  core/proc/downstrike.cc:257-370
  call site                    c_image_stacking_pipeline.cc:763-767 (max_levels = 100, on the result of
                               c_frame_accumulation::compute())

cv::boxFilter is delegated to cv2 (same OpenCV primitive); `box_sum_model` is the scalar model the CUDA kernels
are written from (fp64 sum of the 3x3 BORDER_REPLICATE window rounded once to fp32), pinned against cv2 in
tests/test_inpaint_oracle.py.

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
import cv2

f32 = np.float32


def downstrike_even(src):
    # downstrike.cc:34-78: keeps rows min(2y+1, rows-1) and columns 1, 3, 5 ... clamped to cols-1
    rows, cols = src.shape[:2]
    ys = np.minimum(2 * np.arange((rows + 1) // 2) + 1, rows - 1)
    xs = np.minimum(2 * np.arange((cols + 1) // 2) + 1, cols - 1)
    return np.ascontiguousarray(src[ys][:, xs])


def upject_even(src, dsize):
    # downstrike.cc:257-370: dst(2y+1, 2x+1) = src(y, x) where that fits, zero elsewhere; zmask = 1 on the copied pixels
    cols, rows = dsize
    dst = np.zeros((rows, cols) + src.shape[2:], dtype=src.dtype)
    zmask = np.zeros((rows, cols), dtype=f32)
    ny, nx = len(range(1, rows, 2)), len(range(1, cols, 2))
    dst[1::2, 1::2] = src[:ny, :nx]
    zmask[1::2, 1::2] = 1
    return dst, zmask


def _filter2(src, srcmask, fallback_src, fallback_mask):
    # average_pyramid_inpaint.cc:17-56
    dst = cv2.boxFilter(src, -1, (3, 3), normalize=False, borderType=cv2.BORDER_REPLICATE)
    dstmask = cv2.boxFilter(srcmask, -1, (3, 3), normalize=False, borderType=cv2.BORDER_REPLICATE)
    fb = fallback_mask != 0
    have = ~fb & (dstmask != 0)
    with np.errstate(divide="ignore"):
        scale = (f32(1.0) / dstmask).astype(f32)          # const float scale = 1.0f / *mskp
    if dst.ndim == 3:
        dst[have] = (dst[have] * scale[have][:, None]).astype(f32)
    else:
        dst[have] = (dst[have] * scale[have]).astype(f32)
    dst[fb] = fallback_src[fb]
    dstmask[fb | have] = 1
    return dst, dstmask


def _recurse(image, mask, max_levels):
    # average_pyramid_inpaint.cc:69-95
    if min(image.shape[0], image.shape[1]) > 1 and max_levels > 0:
        fimg, fmsk = _filter2(image, mask, image, mask)
        fimg, fmsk = downstrike_even(fimg), downstrike_even(fmsk)
        if cv2.countNonZero(fmsk) < fmsk.size:
            fimg, fmsk = _recurse(fimg, fmsk, max_levels - 1)
        fimg, fmsk = upject_even(fimg, (image.shape[1], image.shape[0]))
        return _filter2(fimg, fmsk, image, mask)
    return image, mask


def average_pyramid_inpaint(src, mask, max_levels=100):
    """average_pyramid_inpaint(src, mask, dst, dstmask, max_levels) -> (dst, dstmask CV_8U).
    src: CV_32F HxW or HxWxC; mask: CV_8UC1 or None."""
    src = np.ascontiguousarray(src, dtype=f32)
    if mask is None or cv2.countNonZero(mask) == mask.size:
        return src.copy(), (None if mask is None else mask.copy())
    img = np.zeros_like(src)
    m = mask != 0
    img[m] = src[m]                                         # _src.getMat().copyTo(src, mask)
    mskf = _convert_to(mask.astype(np.uint8), cv2.CV_32F, 1.0 / 255.0)    # mask.convertTo(msk, CV_32F, 1.0 / 255.0)
    img, mskf = _recurse(img, mskf, max_levels)
    return img, _convert_to(mskf, cv2.CV_8U, 255.0)                      # msk.convertTo(_dstmask, CV_8U, 255.0)


def _convert_to(a, rtype, alpha):
    # cv::Mat::convertTo has no cv2 binding; OpenCV's cvtScale computes in float for these depths
    # (convert_scale.simd.hpp: 8u->32f (float)src * (float)alpha; 32f->8u saturate_cast<uchar>(src * (float)alpha))
    if rtype == cv2.CV_32F:
        return (a.astype(f32) * f32(alpha)).astype(f32)
    return np.clip(np.rint(a.astype(f32) * f32(alpha)), 0, 255).astype(np.uint8)


def box_sum_model(a):
    """Scalar model of cv::boxFilter(a, CV_32F, Size(3,3), normalize=false, BORDER_REPLICATE): RowSum<float,double> +
    ColumnSum<double,float> = the fp64 sum of the window rounded once to fp32."""
    a = np.asarray(a, dtype=f32)
    p = np.pad(a.astype(np.float64), ((1, 1), (1, 1)) + ((0, 0),) * (a.ndim - 2), mode="edge")
    rows, cols = a.shape[:2]
    s = np.zeros(a.shape, np.float64)
    for dy in range(3):
        r = p[dy:dy + rows, 0:cols] + p[dy:dy + rows, 1:cols + 1] + p[dy:dy + rows, 2:cols + 2]
        s = s + r
    return s.astype(f32)


# ---------------------------------------------------------------------------------------------------------
# linear_interpolation_inpaint (core/proc/inpaint/linear_interpolation_inpaint.cc:14-368): what create_reference_frame
# applies to the generated master frame (c_image_stacking_pipeline.cc:1282-1284) and read_input_frame to frames with a
# missing-pixel mask (c_image_stacking_pipeline_base.cc:258-261).
# ---------------------------------------------------------------------------------------------------------
def _interpolate_holes_1d(image, mask):
    """_interpolate_holes_h2 (:14-117) along axis 1 of image (H x W x C float32) -> (inpaint, dists).
    Every run of holes [start, end) of a row is filled from its neighbours s = start - 1, e = end:
      both sides:   sv + (x - s) * kk with kk = (ev - sv) * (1 / (end - start)),  dist = max(x - s, e - x)
      left only:    sv, dist = x - s;     right only:  ev, dist = e - x;     neither: untouched, dist = 0."""
    h, w = mask.shape
    valid = mask != 0
    idx = np.arange(w, dtype=np.int64)[None, :]
    left = np.maximum.accumulate(np.where(valid, idx, -1), axis=1)               # nearest valid column <= x
    right = np.minimum.accumulate(np.where(valid, idx, w)[:, ::-1], axis=1)[:, ::-1]   # nearest valid column >= x
    hole = ~valid
    has_l, has_r = hole & (left >= 0), hole & (right < w)
    ls, rs = np.clip(left, 0, w - 1), np.clip(right, 0, w - 1)
    rows = np.arange(h)[:, None]
    sv, ev = image[rows, ls], image[rows, rs]
    x = np.broadcast_to(idx, (h, w))
    scale = (f32(1.0) / (right - left - 1).clip(1).astype(f32)).astype(f32)       # 1.0f / (end - start)
    kk = ((ev - sv) * scale[..., None]).astype(f32)
    factor = (x - left).astype(f32)
    both = has_l & has_r
    out = image.copy()
    val_both = (sv + (factor[..., None] * kk).astype(f32)).astype(f32)
    out[both] = val_both[both]
    only_l, only_r = has_l & ~has_r, has_r & ~has_l
    out[only_l] = sv[only_l]
    out[only_r] = ev[only_r]
    dist = np.zeros((h, w), dtype=f32)
    dist[both] = np.maximum(x - left, right - x)[both].astype(f32)
    dist[only_l] = (x - left)[only_l].astype(f32)
    dist[only_r] = (right - x)[only_r].astype(f32)
    return out, dist


def linear_interpolation_inpaint(image, mask):
    """linear_interpolation_inpaint(src, mask, dst) (:327-368) on CV_32F images -> filled copy of `image`."""
    img = np.ascontiguousarray(image, dtype=f32)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[..., None]
    img = img.copy()
    if mask is None:
        return img[..., 0] if squeeze else img
    m = np.ascontiguousarray(mask).copy()
    holes = int(m.size - np.count_nonzero(m))
    while holes > 0:
        ih, dh = _interpolate_holes_1d(img, m)
        iv, dv = _interpolate_holes_1d(np.ascontiguousarray(img.transpose(1, 0, 2)), np.ascontiguousarray(m.T))
        iv, dv = iv.transpose(1, 0, 2), dv.T
        hole = m == 0
        both = hole & (dh > 0) & (dv > 0)
        dd = (f32(1.0) / (dh + dv).clip(1e-30)).astype(f32)
        # dd * (h * dv + v * dh)   (_fill_holes2, :230-309)
        mix = (dd[..., None] * ((ih * dv[..., None]).astype(f32) + (iv * dh[..., None]).astype(f32)).astype(f32)).astype(f32)
        only_v, only_h = hole & ~(dh > 0) & (dv > 0), hole & (dh > 0) & ~(dv > 0)
        img[both] = mix[both]
        img[only_v] = iv[only_v]
        img[only_h] = ih[only_h]
        filled = int(both.sum() + only_v.sum() + only_h.sum())
        m[both | only_v | only_h] = 255
        if filled < 1:
            break
        holes -= filled
    return img[..., 0] if squeeze else img
