"""
Numpy model of the OpenCV primitives whose *exact* semantics the CUDA kernels reproduce
(SURVEY.md Appendix B).  OpenCV is the reference's un-vendored, un-pinned third-party dependency
(CMakeLists.txt:47-59, `find_package(OpenCV)` >= 4.2); its published algorithms for
cv::remap / cv::pyrDown / cv::sepFilter2D are restated here and checked against cv2 4.13.0 in
tests/test_cvmodel.py.  The model is the specification the kernels are written to.

Test infrastructure only (see oracle/__init__.py).
"""
import math
import numpy as np

f32 = np.float32
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS          # 32
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS

BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT101, BORDER_TRANSPARENT = 0, 1, 2, 3, 4, 5


def quantize(coord):
    """cv::remap converts a CV_32FC2 map with cvRound(v * 32) (round-half-even, evaluated in float)."""
    s = np.rint((coord.astype(f32) * f32(INTER_TAB_SIZE)).astype(f32)).astype(np.int64)
    return s >> INTER_BITS, s & (INTER_TAB_SIZE - 1)


def border_interpolate(p, n, border):
    """cv::borderInterpolate for the modes used on the path; returns index or -1 (constant/transparent)."""
    p = np.asarray(p, dtype=np.int64).copy()
    if border == BORDER_REPLICATE:
        return np.clip(p, 0, n - 1)
    if border in (BORDER_REFLECT, BORDER_REFLECT101):
        delta = 1 if border == BORDER_REFLECT101 else 0
        if n == 1:
            return np.zeros_like(p)
        for _ in range(8):
            neg = p < 0
            p[neg] = -p[neg] - 1 + delta
            big = p >= n
            p[big] = n - 1 - (p[big] - n) - delta
        return p
    if border == BORDER_WRAP:
        return np.mod(p, n)
    out = p.copy()
    out[(p < 0) | (p >= n)] = -1
    return out


def linear_coeffs():
    """initInterTab1D(INTER_LINEAR): (1 - x, x) for x = i/32, float."""
    t = np.arange(INTER_TAB_SIZE, dtype=f32) * f32(1.0 / INTER_TAB_SIZE)
    return np.stack([f32(1) - t, t], axis=1).astype(f32)


def cubic_coeffs():
    """interpolateCubic, A = -0.75, evaluated in float at x = i/32."""
    A = f32(-0.75)
    x = np.arange(INTER_TAB_SIZE, dtype=f32) * f32(1.0 / INTER_TAB_SIZE)
    c = np.empty((INTER_TAB_SIZE, 4), dtype=f32)
    c[:, 0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c[:, 1] = ((A + 2) * x - (A + 3)) * x * x + 1
    c[:, 2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    c[:, 3] = f32(1) - c[:, 0] - c[:, 1] - c[:, 2]
    return c


def fixed_tab_2d(tab1d):
    """initInterTab2D(..., fixpt=true): short weights with the sum forced to 32768.
    Returns int array [fy, fx, ky, kx]."""
    k = tab1d.shape[1]
    out = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, k, k), dtype=np.int64)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            v = (tab1d[i][:, None] * tab1d[j][None, :]).astype(f32)
            it = np.clip(np.rint((v * f32(INTER_REMAP_COEF_SCALE)).astype(f32)), -32768, 32767).astype(np.int64)
            isum = int(it.sum())
            if isum != INTER_REMAP_COEF_SCALE:
                diff = isum - INTER_REMAP_COEF_SCALE
                k2 = k // 2
                Mk = mk = (k2, k2)
                for k1 in range(k2, min(k, k2 + 2)):
                    for kk in range(k2, min(k, k2 + 2)):
                        if it[k1, kk] < it[mk]:
                            mk = (k1, kk)
                        elif it[k1, kk] > it[Mk]:
                            Mk = (k1, kk)
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            out[i, j] = it
    return out


def remap_f32(src, rmap, interp="linear", border=BORDER_REPLICATE, border_value=0.0, dst=None):
    """cv::remap for CV_32FC1 source, CV_32FC2 map."""
    h, w = src.shape
    ix, fx = quantize(rmap[..., 0])
    iy, fy = quantize(rmap[..., 1])
    if interp == "nearest":
        ix = np.rint(rmap[..., 0]).astype(np.int64)
        iy = np.rint(rmap[..., 1]).astype(np.int64)
        taps, off, cx, cy = 1, 0, None, None
    elif interp == "linear":
        taps, off = 2, 0
        c = linear_coeffs()
        cx, cy = c[fx], c[fy]
    else:
        taps, off = 4, -1
        c = cubic_coeffs()
        cx, cy = c[fx], c[fy]
    out = np.zeros(rmap.shape[:2], dtype=f32)
    # BORDER_TRANSPARENT: pixels whose anchor tap (ix, iy) is outside keep dst; the other taps are clamped
    bm = BORDER_REPLICATE if border == BORDER_TRANSPARENT else border
    anyin = np.zeros(rmap.shape[:2], dtype=bool)
    for ky in range(taps):
        yy = border_interpolate(iy + off + ky, h, bm)
        row = np.zeros(rmap.shape[:2], dtype=f32)
        for kx in range(taps):
            xx = border_interpolate(ix + off + kx, w, bm)
            ok = (xx >= 0) & (yy >= 0)
            anyin |= ok
            v = np.where(ok, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], f32(border_value)).astype(f32)
            if taps == 1:
                row = v
            elif taps == 2:
                # remapBilinear: ((S00*w00 + S01*w01) + S10*w10) + S11*w11 -- bit-exact against cv2
                term = (v * (cy[..., ky] * cx[..., kx]).astype(f32)).astype(f32)
                out = term if (ky == 0 and kx == 0) else (out + term).astype(f32)
            else:
                wgt = (cy[..., ky] * cx[..., kx]).astype(f32)
                row = (row + v * wgt).astype(f32)
        if taps == 1:
            out = row
        elif taps == 4:
            out = (out + row).astype(f32)
    if border == BORDER_TRANSPARENT and dst is not None:
        outside = (ix < 0) | (ix >= w) | (iy < 0) | (iy >= h)
        out = np.where(outside, dst, out)
    return out


def remap_u8_all255_valid(size, rmap, interp="linear", thresh=255):
    """(cv::remap(Mat1b(size,255), rmap, interp, BORDER_CONSTANT 0) >= thresh) modelled with the
    fixed-point tables: value = saturate_u8((255 * sum_inbounds(w) + 2^14) >> 15)."""
    w, h = size
    ix, fx = quantize(rmap[..., 0])
    iy, fy = quantize(rmap[..., 1])
    if interp == "linear":
        taps, off, tab = 2, 0, fixed_tab_2d(linear_coeffs())
    else:
        taps, off, tab = 4, -1, fixed_tab_2d(cubic_coeffs())
    S = np.zeros(rmap.shape[:2], dtype=np.int64)
    for ky in range(taps):
        yy = iy + off + ky
        for kx in range(taps):
            xx = ix + off + kx
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            S += np.where(ok, tab[fy, fx, ky, kx], 0)
    val = np.clip((255 * S + (1 << 14)) >> 15, 0, 255)
    return val >= thresh, val


def chol_solve_f32(A, B):
    """cv::solve / cv::invert with DECOMP_CHOLESKY on CV_32F data for n > 3: hal::Cholesky32f (CholImpl<float>,
    modules/core/src/matrix_decomp.cpp): float storage, float products, double accumulators, the diagonal of L
    kept as its reciprocal.  Returns (ok, X) with X float32; B may hold several right-hand sides."""
    L = np.array(A, dtype=f32, copy=True)
    b = np.array(B, dtype=f32, copy=True).reshape(L.shape[0], -1)
    m, n = L.shape[0], b.shape[1]
    eps = float(np.finfo(np.float32).eps)
    for i in range(m):
        for j in range(i):
            s = float(L[i, j])
            for k in range(j):
                s -= float(f32(L[i, k] * L[j, k]))
            L[i, j] = f32(s * float(L[j, j]))
        s = float(L[i, i])
        for k in range(i):
            t = float(L[i, k])
            s -= t * t
        if s < eps:
            return False, np.zeros_like(b)
        L[i, i] = f32(1.0 / np.sqrt(s))
    for i in range(m):
        for j in range(n):
            s = float(b[i, j])
            for k in range(i):
                s -= float(f32(L[i, k] * b[k, j]))
            b[i, j] = f32(s * float(L[i, i]))
    for i in range(m - 1, -1, -1):
        for j in range(n):
            s = float(b[i, j])
            for k in range(m - 1, i, -1):
                s -= float(f32(L[k, i] * b[k, j]))
            b[i, j] = f32(s * float(L[i, i]))
    return True, b


def invert_affine_f32(M):
    """cv::invertAffineTransform on CV_32F data as cv2 4.13 evaluates it (bit-exact, tests/test_cvmodel.py)."""
    M = np.asarray(M, dtype=f32).reshape(2, 3)
    D = f32(f32(M[0, 0] * M[1, 1]) - f32(M[0, 1] * M[1, 0]))
    Di = f32(f32(1.0) / D) if D != 0 else f32(0)
    A11, A22 = f32(M[1, 1] * Di), f32(M[0, 0] * Di)
    A12, A21 = f32(-M[0, 1] * Di), f32(-M[1, 0] * Di)
    b1 = f32(-(float(A11) * float(M[0, 2]) + float(A12) * float(M[1, 2])))
    b2 = f32(-(float(A21) * float(M[0, 2]) + float(A22) * float(M[1, 2])))
    return np.array([[A11, A12, b1], [A21, A22, b2]], dtype=f32)


def pyrdown_f32(src, dsize=None):
    """cv::pyrDown(CV_32FC1, BORDER_REFLECT_101) as cv2 4.13 evaluates it (bit-exact, tests/test_cvmodel.py):
    4-lane universal-intrinsic loops without FMA for the bulk, scalar expressions for the left border column,
    the tail columns and the right border (PyrDownInvoker, modules/imgproc/src/pyramids.cpp). dsize = (w, h)."""
    src = np.asarray(src, dtype=f32)
    h, w = src.shape
    dw, dh = dsize if dsize else ((w + 1) // 2, (h + 1) // 2)
    width0 = min((w - 3) // 2 + 1, dw)

    def refl(p, n):
        p = np.asarray(p).copy()
        for _ in range(4):
            p = np.where(p < 0, -p, p)
            p = np.where(p >= n, 2 * (n - 1) - p, p)
        return p

    xs = np.arange(dw)
    col = lambda d: src[:, refl(2 * xs + d, w)]
    c, l1, r1, l2, r2 = col(0), col(-1), col(1), col(-2), col(2)
    a1 = ((l1 + r1).astype(f32) * f32(4)).astype(f32)
    c6 = (c * f32(6)).astype(f32)
    simd = (c6 + (a1 + (l2 + r2).astype(f32)).astype(f32)).astype(f32)
    scal = (((c6 + a1).astype(f32) + l2).astype(f32) + r2).astype(f32)
    use = (xs >= 1) & (xs < 1 + 4 * ((width0 - 1) // 4))
    rows = np.where(use[None, :], simd, scal)
    ys = np.arange(dh)
    rw = lambda d: rows[refl(2 * ys + d, h), :]
    c, u1, d1, u2, d2 = rw(0), rw(-1), rw(1), rw(-2), rw(2)
    a13 = (u1 + d1).astype(f32)
    vs = (((a13 + c).astype(f32) * f32(4)).astype(f32) + ((u2 + d2).astype(f32) + (c + c).astype(f32)).astype(f32)).astype(f32)
    vc = ((((c * f32(6)).astype(f32) + (a13 * f32(4)).astype(f32)).astype(f32) + u2).astype(f32) + d2).astype(f32)
    out = np.where((xs < 4 * (dw // 4))[None, :], vs, vc)
    return (out * f32(1.0 / 256.0)).astype(f32)


def pyrup_f32(src, dsize):
    """cv::pyrUp(CV_32FC1, dstsize=(w, h)) as cv2 4.13 evaluates it (bit-exact, tests/test_cvmodel.py):
    row pass  even column 2x = (s[x-1] + 6 s[x]) + s[x+1] (x = 0: 6 s0 + 2 s1; x = w-1: s[w-2] + 7 s[w-1]),
              odd column 2x+1 = 4 (s[x] + s[x+1]) (x = w-1: 8 s[w-1]); a dst wider than 2w repeats the last column;
    col pass  even row = (6 r[y] + r[y-1]) + r[y+1], odd row = 4 (r[y] + r[y+1]) with r[-1] = r[1], r[h] = r[h-1];
    the result is scaled by 1/64 (PyrUpInvoker, modules/imgproc/src/pyramids.cpp)."""
    s = np.asarray(src, dtype=f32)
    h, w = s.shape
    dw, dh = dsize
    xs = np.arange(w)
    left = s[:, np.where(xs - 1 < 0, 1, xs - 1)]
    right = s[:, np.minimum(xs + 1, w - 1)]
    t0 = ((left + (s * f32(6)).astype(f32)).astype(f32) + right).astype(f32)
    t1 = ((s + right).astype(f32) * f32(4)).astype(f32)
    t0[:, 0] = ((s[:, 0] * f32(6)).astype(f32) + (s[:, 1] * f32(2)).astype(f32)).astype(f32)
    t0[:, w - 1] = (s[:, w - 2] + (s[:, w - 1] * f32(7)).astype(f32)).astype(f32)
    t1[:, w - 1] = (s[:, w - 1] * f32(8)).astype(f32)
    R = np.zeros((h, 2 * w), f32)
    R[:, 0::2] = t0
    R[:, 1::2] = t1
    if dw > 2 * w:
        R = np.concatenate([R, R[:, -1:]], 1)
    R = R[:, :dw]
    ys = np.arange(h)
    up = R[np.where(ys - 1 < 0, 1, ys - 1)]
    dn = R[np.minimum(ys + 1, h - 1)]
    d0 = (((R * f32(6)).astype(f32) + up).astype(f32) + dn).astype(f32)
    d1 = ((R + dn).astype(f32) * f32(4)).astype(f32)
    D = np.zeros((2 * h, dw), f32)
    D[0::2] = d0
    D[1::2] = d1
    if dh > 2 * h:
        D = np.concatenate([D, D[-1:]], 0)
    return (D[:dh] * f32(1.0 / 64.0)).astype(f32)


# ---------------------------------------------------------------------------------------------------------
# cv::resize(INTER_AREA), down-scaling, CV_32FC1 (modules/imgproc/src/resize.cpp: computeResizeAreaTab,
# ResizeArea_Invoker, ResizeAreaFast_Invoker, ResizeAreaFastVec_SIMD_32f).  Used by scaleImage for ecc.scale other
# than 0.5 (c_frame_registration.cc:242-247) and by compute_local_variance_map's uscale stage
# (c_local_variance_sharpness_measure.cc:231-234).
# ---------------------------------------------------------------------------------------------------------
def _area_tab(ssize, dsize, scale):
    import math
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        cell = min(scale, ssize - fsx1)
        fsx2 = fsx1 + cell
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_f32(src, dsize, inv_scale=None, simd_lanes=8):
    """cv2.resize(src, dsize, interpolation=INTER_AREA) for scale >= 1 (pure Python loops: small images only).
    inv_scale = (fx, fy) when the call derived dsize from them (the scale is then 1/fx, not ssize/dsize).
    simd_lanes: vector width of the 2x2 fast path of the OpenCV build (8 = AVX2)."""
    s = np.asarray(src, dtype=f32)
    sh, sw = s.shape
    dw, dh = dsize
    ix, iy = inv_scale if inv_scale is not None else (dw / sw, dh / sh)
    scale_x, scale_y = 1.0 / ix, 1.0 / iy
    isx, isy = int(round(scale_x)), int(round(scale_y))
    dst = np.zeros((dh, dw), f32)
    eps = np.finfo(np.float64).eps
    if abs(scale_x - isx) < eps and abs(scale_y - isy) < eps:
        wfull = min(dw, sw // isx)
        sc = f32(1.0 / (isx * isy))
        for dy in range(dh):
            y0 = dy * isy
            for dx in range(dw):
                x0 = dx * isx
                if x0 + isx > sw or y0 + isy > sh:      # cell clipped by the image edge: running sum / count
                    vals = [s[y0 + ky, x0 + kx] for ky in range(isy) if y0 + ky < sh for kx in range(isx) if x0 + kx < sw]
                    acc = f32(0)
                    for v in vals:
                        acc = f32(acc + v)
                    dst[dy, dx] = f32(acc / f32(len(vals))) if vals else 0
                    continue
                vals = [s[y0 + ky, x0 + kx] for ky in range(isy) for kx in range(isx)]
                if isx == 2 and isy == 2 and dx < wfull - wfull % simd_lanes:
                    acc = f32(f32(vals[0] + vals[1]) + f32(vals[2] + vals[3]))
                else:
                    acc, k = f32(0), 0
                    while k + 4 <= len(vals):
                        acc = f32(acc + f32(f32(f32(vals[k] + vals[k + 1]) + vals[k + 2]) + vals[k + 3]))
                        k += 4
                    while k < len(vals):
                        acc = f32(acc + vals[k])
                        k += 1
                dst[dy, dx] = f32(acc * sc)
        return dst
    xtab, ytab = _area_tab(sw, dw, scale_x), _area_tab(sh, dh, scale_y)
    prev_dy = ytab[0][0]
    summ = np.zeros(dw, f32)
    for dy, sy, beta in ytab:
        buf = np.zeros(dw, f32)
        for dx, sx, alpha in xtab:
            buf[dx] = f32(buf[dx] + f32(s[sy, sx] * alpha))
        if dy != prev_dy:
            dst[prev_dy] = summ
            summ = (beta * buf).astype(f32)
            prev_dy = dy
        else:
            summ = (summ + (beta * buf).astype(f32)).astype(f32)
    dst[prev_dy] = summ
    return dst


def remap_linear_u8(mask, mapx, mapy):
    """cv::remap of a CV_8UC1 image with INTER_LINEAR and BORDER_CONSTANT 0 (the mask remaps of ecc2.cc:205-216, 1311, 117):
    coordinates quantised to 1/32 px, 15-bit fixed-point bilinear weights - for 1/32 fractions exactly
    32 (32 - fx | fx)(32 - fy | fy), sum 32768 - and FixedPtCast's (sum + 2^14) >> 15."""
    h, w = mask.shape
    sx = np.rint(mapx.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(mapy.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy, fx, fy = sx >> 5, sy >> 5, sx & 31, sy & 31

    def at(x, y):
        ok = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        return np.where(ok, mask[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)], 0).astype(np.int64)

    s = (32 - fx) * (32 - fy) * at(ix, iy) + fx * (32 - fy) * at(ix + 1, iy) + (32 - fx) * fy * at(ix, iy + 1) + fx * fy * at(ix + 1, iy + 1)
    return ((32 * s + (1 << 14)) >> 15).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------
# cv::Mat::dot on CV_32F data WITHOUT IPP (OpenCV modules/core/src/matmul.simd.hpp, dotProd_32f): what a distribution
# build of OpenCV (the reference links the system library, CMakeLists.txt:47-59) computes for H = <J_i, J_j> and
# v = <J_i, rhs> (ecc2.cc:295-338).  The opencv-python wheel in this image has IPP, whose ippsDotProd_32f64f accumulates in
# double, and no Python entry point reaches Mat::dot (cv2.UMat exposes no dot; cv2.gemm / mulTransposed / norm take other
# routines), so this model cannot be pinned to a cv2 call: it restates the published source.
#   blocks of 2^13 elements; inside a block four vector accumulators of `lanes` floats updated by fused multiply-add
#   (v_muladd), combined as v_sum + ((v_sum1 + v_sum2) + v_sum3), left-over vectors into v_sum, horizontal sum in float
#   (halves folded, then adjacent pairs), block sums added in double; scalar tail in double.
# ---------------------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    """fp32 fused multiply-add: the product of two floats is exact in double; one rounding to double (negligibly often
    inexact), one to float."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _reduce_sum_f32(v):
    """v_reduce_sum of a float vector (..., lanes): fold the halves until 4 lanes remain, then adjacent pairs."""
    f = np.float32
    while v.shape[-1] > 4:
        h = v.shape[-1] // 2
        v = (v[..., :h] + v[..., h:]).astype(f)
    v = (v[..., 0::2] + v[..., 1::2]).astype(f)
    return (v[..., 0] + v[..., 1]).astype(f)


def dot_f32_simd(a, b, lanes=16):
    """dotProd_32f(a, b) of OpenCV's universal-intrinsics path (lanes = 4 SSE/NEON, 8 AVX2, 16 AVX-512) -> double."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1)
    n = a.size
    len0 = n & -(lanes * 4)
    bs0 = 1 << 13
    r = 0.0
    i = 0
    nfull = len0 // bs0
    blocks = []
    if nfull:
        blocks.append((a[:nfull * bs0].reshape(nfull, bs0), b[:nfull * bs0].reshape(nfull, bs0)))
        i = nfull * bs0
    if len0 > i:
        blocks.append((a[i:len0].reshape(1, -1), b[i:len0].reshape(1, -1)))
        i = len0
    for A, B in blocks:
        nb, bs = A.shape
        w = lanes * 4
        nun = bs // w
        acc = np.zeros((nb, 4, lanes), np.float32)
        Au = A[:, :nun * w].reshape(nb, nun, 4, lanes)
        Bu = B[:, :nun * w].reshape(nb, nun, 4, lanes)
        for j in range(nun):
            acc = _fma32(Au[:, j], Bu[:, j], acc)
        vsum = (acc[:, 0] + ((acc[:, 1] + acc[:, 2]).astype(np.float32) + acc[:, 3]).astype(np.float32)).astype(np.float32)
        for j in range(nun * w, bs - lanes + 1, lanes):
            vsum = _fma32(A[:, j:j + lanes], B[:, j:j + lanes], vsum)
        r += float(_reduce_sum_f32(vsum).astype(np.float64).sum())
    if i < n:
        r += float(np.dot(a[i:].astype(np.float64), b[i:].astype(np.float64)))
    return r


# ---------------------------------------------------------------------------------------------------------
# cv::remap(INTER_LANCZOS4) (OpenCV imgproc/imgwarp.cpp: interpolateLanczos4, initInterTab1D / initInterTab2D, remapLanczos4)
# ---------------------------------------------------------------------------------------------------------
def lanczos4_tab():
    """[32][8] float coefficients of interpolateLanczos4 at f / 32: sin / cos in double, coefficients and their sum in float."""
    s45 = 0.70710678118654752440084436210485
    cs = [[1, 0], [-s45, -s45], [0, 1], [s45, -s45], [-1, 0], [s45, s45], [0, -1], [-s45, s45]]
    tab = np.zeros((32, 8), f32)
    for i in range(32):
        x = f32(f32(i) * f32(1.0 / 32))
        c = np.zeros(8, f32)
        ssum = f32(0)
        y0 = -(float(x) + 3) * math.pi * 0.25
        s0, c0 = math.sin(y0), math.cos(y0)
        for k in range(8):
            y0_ = f32(x + f32(3) - f32(k))
            if abs(float(y0_)) >= 1e-6:
                y = -float(y0_) * math.pi * 0.25
                c[k] = f32((cs[k][0] * s0 + cs[k][1] * c0) / (y * y))
            else:
                c[k] = f32(1e30)
            ssum = f32(ssum + c[k])
        tab[i] = (c * f32(f32(1) / ssum)).astype(f32)
    return tab


def lanczos4_itab():
    """[32][32][8][8] fixed-point 2-D weights (initInterTab2D, fixpt): cvRound(w * 32768) with the sum forced to 32768 on the
    largest / smallest of the four central entries."""
    tab = lanczos4_tab()
    out = np.zeros((32, 32, 8, 8), np.int32)
    for i in range(32):
        for j in range(32):
            w = (tab[i][:, None] * tab[j][None, :]).astype(f32)
            it = np.clip(np.rint((w * f32(32768)).astype(f32)), -32768, 32767).astype(np.int32)
            diff = int(it.sum()) - 32768
            if diff != 0:
                Mk, mk = (4, 4), (4, 4)
                for k1 in (4, 5):
                    for k2 in (4, 5):
                        if it[k1, k2] < it[mk]:
                            mk = (k1, k2)
                        elif it[k1, k2] > it[Mk]:
                            Mk = (k1, k2)
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            out[i, j] = it
    return out


def remap_lanczos4_f32(src, mapx, mapy):
    """cv::remap(src CV_32FC1, INTER_LANCZOS4) where the 8 x 8 footprint lies inside the image (NaN elsewhere): the float weight
    of tap (ky, kx) is tab[fy][ky] * tab[fx][kx]; a row's eight products are added left to right, the rows top to bottom."""
    tab = lanczos4_tab()
    h, w = src.shape
    sx = np.rint(mapx.astype(f32) * f32(32)).astype(np.int64)
    sy = np.rint(mapy.astype(f32) * f32(32)).astype(np.int64)
    ix, iy, fx, fy = (sx >> 5) - 3, (sy >> 5) - 3, sx & 31, sy & 31
    inside = (ix >= 0) & (ix + 8 <= w) & (iy >= 0) & (iy + 8 <= h)
    ixc, iyc = np.clip(ix, 0, w - 8), np.clip(iy, 0, h - 8)
    out = np.zeros(mapx.shape, f32)
    for r in range(8):
        wy = tab[fy, r]
        row = None
        for k in range(8):
            term = (src[iyc + r, ixc + k] * (wy * tab[fx, k]).astype(f32)).astype(f32)
            row = term if row is None else (row + term).astype(f32)
        out = (out + row).astype(f32)
    out[~inside] = np.nan
    return out


def remap_lanczos4_all255_valid(size, mapx, mapy):
    """(cv::remap(Mat1b(size, 255), INTER_LANCZOS4, BORDER_CONSTANT 0) >= 255) through the fixed-point table."""
    h, w = size
    itab = lanczos4_itab()
    sx = np.rint(mapx.astype(f32) * f32(32)).astype(np.int64)
    sy = np.rint(mapy.astype(f32) * f32(32)).astype(np.int64)
    ix, iy, fx, fy = (sx >> 5) - 3, (sy >> 5) - 3, sx & 31, sy & 31
    S = np.zeros(mapx.shape, np.int64)
    for r in range(8):
        oky = (iy + r >= 0) & (iy + r < h)
        for k in range(8):
            ok = oky & (ix + k >= 0) & (ix + k < w)
            S += np.where(ok, itab[fy, fx, r, k], 0)
    return np.clip((255 * S + (1 << 14)) >> 15, 0, 255) >= 255


# ---------------------------------------------------------------------------------------------------------
# cv::resize on CV_32F data as a non-IPP OpenCV build computes it (imgproc/resize.cpp): the per-axis tables the device code of
# c_eccflow (ssk_eccflow.cu) and of the frame up-scaling (ssk_upscale.cuh) is written from
# ---------------------------------------------------------------------------------------------------------
def resize_cubic_tab(ssize, dsize):
    """INTER_CUBIC: fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), interpolateCubic(fx - s) (A = -0.75) in float.
    -> (s [dsize], coeffs [dsize][4]); the taps are s - 1 .. s + 2, clamped into the image."""
    scale = 1.0 / (dsize / ssize)
    s = np.zeros(dsize, np.int64)
    c = np.zeros((dsize, 4), f32)
    A = f32(-0.75)
    for d in range(dsize):
        fx = f32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = f32(fx - f32(sx))
        c0 = f32(f32(f32(f32(f32(f32(A * f32(fx + 1)) - f32(5 * A)) * f32(fx + 1)) + f32(8 * A)) * f32(fx + 1)) - f32(4 * A))
        c1 = f32(f32(f32(f32(f32(f32(A + 2) * fx) - f32(A + 3)) * fx) * fx) + f32(1))
        g = f32(1 - fx)
        c2 = f32(f32(f32(f32(f32(f32(A + 2) * g) - f32(A + 3)) * g) * g) + f32(1))
        c3 = f32(f32(f32(f32(1) - c0) - c1) - c2)
        s[d] = sx
        c[d] = [c0, c1, c2, c3]
    return s, c


def resize_cubic_f32(src, dsize):
    """cv2.resize(src, dsize, interpolation=INTER_CUBIC) on CV_32F (1 or more channels): horizontal pass, then vertical."""
    sh, sw = src.shape[:2]
    dw, dh = dsize
    xs, xc = resize_cubic_tab(sw, dw)
    ys, yc = resize_cubic_tab(sh, dh)
    H = np.zeros((sh, dw) + src.shape[2:], f32)
    for dx in range(dw):
        idx = np.clip(xs[dx] - 1 + np.arange(4), 0, sw - 1)
        v = (src[:, idx[0]] * xc[dx, 0]).astype(f32)
        for k in range(1, 4):
            v = (v + (src[:, idx[k]] * xc[dx, k]).astype(f32)).astype(f32)
        H[:, dx] = v
    out = np.zeros((dh, dw) + src.shape[2:], f32)
    for dy in range(dh):
        idx = np.clip(ys[dy] - 1 + np.arange(4), 0, sh - 1)
        v = (H[idx[0]] * yc[dy, 0]).astype(f32)
        for k in range(1, 4):
            v = (v + (H[idx[k]] * yc[dy, k]).astype(f32)).astype(f32)
        out[dy] = v
    return out


def resize_area_tab(ssize, dsize):
    """computeResizeAreaTab for one axis -> per destination index the list of (source index, float weight)."""
    scale = 1.0 / (dsize / ssize)
    out = []
    for dx in range(dsize):
        f1 = dx * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        e = []
        if s1 - f1 > 1e-3:
            e.append((s1 - 1, f32((s1 - f1) / cell)))
        for sx in range(s1, s2):
            e.append((sx, f32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            e.append((s2, f32(min(min(f2 - s2, 1.0), cell) / cell)))
        out.append(e)
    return out


def resize_area_tables_f32(src, dsize):
    """INTER_AREA through the per-axis tables, columns first (the order ssk_eccflow.cu::k_flow_reduce sums in); equals
    cv2.resize(INTER_AREA) to float rounding for any scale >= 1 (cv2 sums rows first and takes a 1 / area path for integer scales)."""
    sh, sw = src.shape[:2]
    dw, dh = dsize
    xt, yt = resize_area_tab(sw, dw), resize_area_tab(sh, dh)
    col = np.zeros((dh, sw) + src.shape[2:], np.float64)
    for dy, e in enumerate(yt):
        for sy, b in e:
            col[dy] += float(b) * src[sy]
    out = np.zeros((dh, dw) + src.shape[2:], np.float64)
    for dx, e in enumerate(xt):
        for sx, a in e:
            out[:, dx] += float(a) * col[:, sx]
    return out.astype(f32)


def resize_linear_axis(ssize, dsize):
    """INTER_LINEAR: (s0, s1, a0, a1) per destination index; the fraction is zeroed where the taps would leave the image."""
    scale = 1.0 / (dsize / ssize)
    s0, s1 = np.zeros(dsize, np.int64), np.zeros(dsize, np.int64)
    a0, a1 = np.zeros(dsize, f32), np.zeros(dsize, f32)
    for d in range(dsize):
        fx = f32((d + 0.5) * scale - 0.5)
        s = int(np.floor(fx))
        fx = f32(fx - f32(s))
        if s < 0:
            fx, s = f32(0), 0
        if s >= ssize - 1:
            fx, s = f32(0), ssize - 1
        s0[d], s1[d], a0[d], a1[d] = s, min(s + 1, ssize - 1), f32(1) - fx, fx
    return s0, s1, a0, a1


def resize_linear_f32(src, dsize):
    """cv2.resize(src, dsize, interpolation=INTER_LINEAR) on CV_32FC1 without IPP: S0 a0 + S1 a1 per axis, no fused multiply-add."""
    sh, sw = src.shape
    dw, dh = dsize
    x0, x1, ax0, ax1 = resize_linear_axis(sw, dw)
    y0, y1, ay0, ay1 = resize_linear_axis(sh, dh)
    H = ((src[:, x0] * ax0).astype(f32) + (src[:, x1] * ax1).astype(f32)).astype(f32)
    return ((H[y0] * ay0[:, None]).astype(f32) + (H[y1] * ay1[:, None]).astype(f32)).astype(f32)


def resize_linear_u8_ge255(mask, dsize):
    """(cv2.resize(mask CV_8UC1, dsize, INTER_LINEAR) >= 255) without IPP: 11-bit fixed-point coefficients cvRound(f * 2048),
    horizontal pass in int, vertical ((b0 (S0 >> 4)) >> 16) + ((b1 (S1 >> 4)) >> 16) + 2) >> 2."""
    sh, sw = mask.shape
    dw, dh = dsize
    x0, x1, ax0, ax1 = resize_linear_axis(sw, dw)
    y0, y1, ay0, ay1 = resize_linear_axis(sh, dh)
    ia0, ia1 = np.rint(ax0 * f32(2048)).astype(np.int64), np.rint(ax1 * f32(2048)).astype(np.int64)
    ib0, ib1 = np.rint(ay0 * f32(2048)).astype(np.int64), np.rint(ay1 * f32(2048)).astype(np.int64)
    m = mask.astype(np.int64)
    H = m[:, x0] * ia0 + m[:, x1] * ia1
    v = (((ib0[:, None] * (H[y0] >> 4)) >> 16) + ((ib1[:, None] * (H[y1] >> 4)) >> 16) + 2) >> 2
    return v >= 255
