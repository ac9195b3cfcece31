"""
Oracle restatement of the per-frame weight-map generators.

W1  compute_local_variance_map   core/proc/sharpness_measure/c_local_variance_sharpness_measure.cc:15-247
    (called from c_image_stacking_pipeline::compute_weights, c_image_stacking_pipeline.cc:2013-2020;
     option defaults dscale=1, kradius=1, uscale=0: c_image_stacking_pipeline.h:94-99)
W2  lpg                          core/proc/lpg.cc:60-129 (5x5 stencil), :223-290 (driver)

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np
import cv2

f32 = np.float32


def _maxval(dtype):
    # pixtype.cc:93-108 / lpg.cc:184-200
    return {np.dtype(np.uint8): 255.0, np.dtype(np.int8): 127.0, np.dtype(np.uint16): 65535.0,
            np.dtype(np.int16): 32767.0, np.dtype(np.int32): 2147483647.0}.get(np.dtype(dtype), 1.0)


def _pdownscale(src, level, border_mode=cv2.BORDER_DEFAULT):
    # c_local_variance_sharpness_measure.cc:28-52 and lpg.cc:132-156 (same code)
    if min(src.shape[0], src.shape[1]) < 4:
        return src.copy()
    dst = cv2.pyrDown(src, borderType=border_mode)
    if min(dst.shape[0], dst.shape[1]) >= 4:
        for _ in range(1, level):
            dst = cv2.pyrDown(dst, borderType=border_mode)
            if min(dst.shape[0], dst.shape[1]) < 4:
                break
    return dst


def _dscale_size(size, level):
    # c_local_variance_sharpness_measure.cc:15-25 ; size=(w,h)
    for _ in range(level):
        nxt = ((size[0] + 1) // 2, (size[1] + 1) // 2)
        if min(nxt) < 4:
            break
        size = nxt
    return size


def _pupscale(image, dst_size):
    # lpg.cc:158-182 ; dst_size=(w,h)
    h, w = image.shape[:2]
    if (w, h) == tuple(dst_size):
        return image
    sizes = [tuple(dst_size)]
    while True:
        nxt = ((sizes[-1][0] + 1) // 2, (sizes[-1][1] + 1) // 2)
        if nxt == (w, h):
            break
        if nxt[0] < w or nxt[1] < h:
            raise RuntimeError("invalid next size")
        sizes.append(nxt)
    for s in reversed(sizes):
        image = cv2.pyrUp(image, dstsize=s)
    return image


def compute_local_variance_map(image, dscale=1, kradius=1, uscale=0, full_resolution=True):
    """c_local_variance_sharpness_measure.cc:193-247 -> (Q, map or None). Single-channel input
    (colour goes through extract_channel(gray) first, same cvtColor as registration.create_ecc_image)."""
    ksize = 2 * max(1, kradius) + 1
    SE = np.full((ksize, ksize), 255, dtype=np.uint8)
    depth_scale = 20.0 * 1.0 / _maxval(image.dtype)
    M = image if image.ndim == 2 else cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    if dscale > 0:
        M = _pdownscale(M, dscale)
    G = cv2.morphologyEx(M, cv2.MORPH_GRADIENT, SE, borderType=cv2.BORDER_REPLICATE)
    W = cv2.norm(G, cv2.NORM_L1)
    if not (W > 0):
        return 0.0, None
    # _compute_sharpness_map, :122-163 (float arithmetic; the reference's float atomic sum is order dependent,
    # here the g^4 terms are summed in float64 and narrowed once)
    Gf = G.astype(f32)
    map_scale = f32(depth_scale * depth_scale * depth_scale)
    Mm = (Gf * Gf * Gf * map_scale).astype(f32)
    total = f32(np.sum((Gf * Gf * Gf * Gf).astype(f32), dtype=np.float64))
    Q = (depth_scale ** 3 / W) * float(total)
    if uscale > 0:
        Mm = cv2.resize(Mm, _dscale_size((Mm.shape[1], Mm.shape[0]), uscale), interpolation=cv2.INTER_AREA)
    Mm = cv2.add(Mm, (0.05 * Q, 0, 0, 0))
    if full_resolution and Mm.shape[:2] != image.shape[:2]:
        Mm = cv2.resize(Mm, (image.shape[1], image.shape[0]), interpolation=cv2.INTER_LINEAR)
    return Q, Mm


def compute_lpg_5x5(src, alpha, beta, eps):
    """lpg.cc:60-129."""
    h, w = src.shape
    dst = np.zeros((h, w), dtype=f32)
    l_norm = f32(100.0 / 4.0)
    g_norm = f32(100.0 / 36.0)
    alpha = f32(f32(alpha) * (l_norm * l_norm))
    beta = f32(f32(beta) * (g_norm * g_norm))
    eps = f32(eps)

    def r(dy, dx):
        return src[2 + dy:h - 2 + dy, 2 + dx:w - 2 + dx]

    two, four = f32(2), f32(4)
    gx = ((r(-2, 2) + two * r(-1, 2) + four * r(0, 2) + two * r(1, 2) + r(2, 2)) -
          (r(-2, -2) + two * r(-1, -2) + four * r(0, -2) + two * r(1, -2) + r(2, -2)) +
          two * ((r(-1, 1) + two * r(0, 1) + r(1, 1)) - (r(-1, -1) + two * r(0, -1) + r(1, -1))))
    gy = ((r(2, -2) + two * r(2, -1) + four * r(2, 0) + two * r(2, 1) + r(2, 2)) -
          (r(-2, -2) + two * r(-2, -1) + four * r(-2, 0) + two * r(-2, 1) + r(-2, 2)) +
          two * ((r(1, -1) + two * r(1, 0) + r(1, 1)) - (r(-1, -1) + two * r(-1, 0) + r(-1, 1))))
    grad = gx * gx + gy * gy
    lap = (f32(16) * r(0, 0) - two * (r(-1, 0) + r(1, 0) + r(0, -1) + r(0, 1)) -
           (r(-1, -1) + r(-1, 1) + r(1, -1) + r(1, 1)) - (r(-2, 0) + r(2, 0) + r(0, -2) + r(0, 2)))
    lapl = lap * lap
    dst[2:h - 2, 2:w - 2] = (alpha * lapl + beta * grad + eps).astype(f32)
    # row edges: out[0] = out[1] = out[2]; out[cols-1] = out[cols-2] = out[cols-3]
    dst[2:h - 2, 0] = dst[2:h - 2, 2]
    dst[2:h - 2, 1] = dst[2:h - 2, 2]
    dst[2:h - 2, w - 1] = dst[2:h - 2, w - 3]
    dst[2:h - 2, w - 2] = dst[2:h - 2, w - 3]
    if h > 4:
        dst[0] = dst[2]
        dst[1] = dst[2]
        dst[h - 2] = dst[h - 3]
        dst[h - 1] = dst[h - 3]
    return dst


def lpg(image, k=2.0, p=2.0, dscale=2, uscale=6):
    """lpg.cc:223-290 -> map (float32, size of image)."""
    if image.dtype == np.float32:
        s = image.copy()
    else:
        s = (image.astype(np.float64) * (1.0 / _maxval(image.dtype))).astype(f32)
    if s.ndim == 3 and s.shape[2] > 1:
        s = cv2.reduce(s.reshape(-1, s.shape[2]), 1, cv2.REDUCE_AVG).reshape(s.shape[:2])
    if dscale > 0:
        s = _pdownscale(s, dscale)
        s = cv2.multiply(s, (1.0 / (1 + dscale), 0, 0, 0))
    m = compute_lpg_5x5(s, k / (k + 1), 1.0 / (k + 1), 1e-9)
    if uscale > 0 and uscale > dscale:
        m = _pdownscale(m, uscale - dscale)
        m = cv2.multiply(m, (float(uscale - dscale), 0, 0, 0))
    if p != 0 and p != 1:
        m = cv2.pow(m, p)
    if m.shape[:2] != image.shape[:2]:
        m = _pupscale(m, (image.shape[1], image.shape[0]))
    return m
