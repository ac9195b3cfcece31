"""
Oracle restatement of the ECC image alignment (single-level solvers + coarse-to-fine driver).

Follows /root/reference/core/proc/image_registration/ecc2.cc:
  compute_correlation            ecc2.cc:65-137
  ecc_differentiate              ecc2.cc:142-169 (member copy :1838-1863)
  ecc_remap                      ecc2.cc:172-219
  ecc_compute_hessian_matrix     ecc2.cc:295-322
  ecc_project_error_image        ecc2.cc:327-338
  ecc_convert_input_image        ecc2.cc:345-382
  ecc_normalize                  ecc2.cc:385-397
  c_ecc_align base               ecc2.cc:505-690
  c_ecch                         ecc2.cc:944-1176, ecc2.h:290-293
  c_ecc_forward_additive         ecc2.cc:1187-1365
  c_ecclm                        ecc2.cc:1371-1650
  c_ecc_inverse_compositional    ecc2.cc:1656-1788
  c_ecclm_inverse_compositional  ecc2.cc:1794-2086

Every OpenCV primitive the reference calls is called here through cv2 with the same arguments.
Test infrastructure only (see oracle/__init__.py).
"""
import math
import numpy as np
import cv2

f32 = np.float32
FLT_MAX = float(np.finfo(np.float32).max)

# ecc2.h:59-64
ECC_ALIGN_FORWARD_ADDITIVE = 0
ECC_ALIGN_INVERSE_COMPOSITIONAL = 1
ECC_ALIGN_LM = 2
ECC_ALIGN_INVERSE_COMPOSITIONAL_LM = 3

_D5 = np.array([1.0 / 12.0, -2.0 / 3.0, 0.0, 2.0 / 3.0, -1.0 / 12.0], dtype=f32).reshape(-1, 1)
_S3 = np.array([0.25, 0.5, 0.25], dtype=f32).reshape(-1, 1)
_SE5 = np.full((5, 5), 255, dtype=np.uint8)


def _dot(a, b):
    """cv::Mat::dot on CV_32F: float products, accumulated in double."""
    return float(np.dot(a.reshape(-1).astype(np.float64), b.reshape(-1).astype(np.float64)))


def _ddiv(a, b):
    """a / b as C++ computes it on doubles: x / 0 is +-inf, 0 / 0 is NaN (the reference divides by CMA == 0 when a diverging trial
    maps every pixel outside the image, ecc2.cc:1915, 1923)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(a) / np.float64(b))


def _norm_l2sqr(a):
    return float(np.dot(a.reshape(-1).astype(np.float64), a.reshape(-1).astype(np.float64)))


def compute_correlation_masked(src1, src2, mask):
    """ecc2.cc:65-100 (single channel)."""
    npix = src1.size if mask is None else cv2.countNonZero(mask)
    m1, s1 = cv2.meanStdDev(src1, mask=mask)
    m2, s2 = cv2.meanStdDev(src2, mask=mask)
    img1 = cv2.subtract(src1, (float(m1[0, 0]), 0, 0, 0), mask=mask, dtype=cv2.CV_32F)
    img2 = cv2.subtract(src2, (float(m2[0, 0]), 0, 0, 0), mask=mask, dtype=cv2.CV_32F)
    if mask is not None:
        img1[mask == 0] = 0
        img2[mask == 0] = 0
    covar = _dot(img1, img2) / npix
    return covar / (float(s1[0, 0]) * float(s2[0, 0]))


def compute_correlation(current_image, current_mask, reference_image, reference_mask, rmap):
    """ecc2.cc:103-137."""
    remapped_image = cv2.remap(current_image, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    if current_mask is None:
        src_mask = np.full(current_image.shape[:2], 255, dtype=np.uint8)
    else:
        src_mask = current_mask
    remapped_mask = cv2.remap(src_mask, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    remapped_mask = cv2.compare(remapped_mask, 254, cv2.CMP_GE)
    if reference_mask is not None:
        remapped_mask = cv2.bitwise_and(reference_mask, remapped_mask)
    return compute_correlation_masked(reference_image, remapped_image, remapped_mask)


def ecc_differentiate(src, mask=None, inverted=False):
    """ecc2.cc:142-169 (mask = valid mask) / :1838-1863 (inverted=True: mask marks bad pixels)."""
    gx = cv2.sepFilter2D(src, cv2.CV_32F, _D5, _S3, borderType=cv2.BORDER_REPLICATE)
    gy = cv2.sepFilter2D(src, cv2.CV_32F, _S3, _D5, borderType=cv2.BORDER_REPLICATE)
    if mask is not None:
        bad = (mask != 0) if inverted else (mask == 0)
        gx[bad] = 0
        gy[bad] = 0
    return gx, gy


def ecc_remap(transform, params, size, src, src_mask, border=cv2.BORDER_REPLICATE):
    """ecc2.cc:178-219: returns (dst, dst_mask(0/255))."""
    rmap = transform.create_remap(size, params)
    dst = cv2.remap(src, rmap, None, cv2.INTER_LINEAR, borderMode=border)
    m = src_mask if src_mask is not None else np.full(src.shape[:2], 255, dtype=np.uint8)
    dst_mask = cv2.remap(m, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    dst_mask = cv2.compare(dst_mask, 250, cv2.CMP_GE)
    return dst, dst_mask


def ecc_compute_hessian_matrix(J):
    """ecc2.cc:295-322."""
    M = len(J)
    H = np.zeros((M, M), dtype=f32)
    for i in range(M):
        for j in range(i + 1):
            H[i, j] = _dot(J[i], J[j])
    for i in range(M):
        for j in range(i + 1, M):
            H[i, j] = H[j, i]
    return H


def ecc_project_error_image(J, rhs):
    """ecc2.cc:327-338."""
    return np.array([[_dot(Ji, rhs)] for Ji in J], dtype=f32)


def ecc_convert_input_image(src, src_mask):
    """ecc2.cc:345-382."""
    if src_mask is not None:
        assert src_mask.shape[:2] == src.shape[:2] and src_mask.dtype == np.uint8 and src_mask.ndim == 2
    if src.ndim == 2 or src.shape[2] == 1:
        dst = src.reshape(src.shape[:2]).astype(f32)
    else:
        tmp = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
        dst = tmp.astype(f32)
    return dst, (None if src_mask is None else src_mask.copy())


def ecc_downscale(src, level, border_mode):
    """ecc2.cc:224-232."""
    dst = cv2.pyrDown(src, borderType=border_mode)
    for _ in range(1, level):
        dst = cv2.pyrDown(dst, borderType=border_mode)
    return dst


def ecc_upscale(image, dst_size):
    """ecc2.cc:237-272; dst_size = (w, h)."""
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


def ecc_normalize(src, src_mask, lvl):
    """ecc2.cc:385-397."""
    h, w = src.shape[:2]
    m = ecc_downscale(src, lvl, cv2.BORDER_REPLICATE)
    m = ecc_upscale(m, (w, h))
    dst = cv2.subtract(src, m, dtype=cv2.CV_32F)
    if src_mask is not None:
        dst[src_mask == 0] = 0
    return dst


def compute_next_pyramid_layer_size(size):
    """ecc2.h:290-293; size = (w, h)."""
    return (((size[0] + 1) >> 1) & ~1, ((size[1] + 1) >> 1) & ~1)


def _size(img):
    return (img.shape[1], img.shape[0])


class EccAlign:
    """c_ecc_align (ecc2.h:70-150, ecc2.cc:505-690)."""

    def __init__(self, transform=None):
        self.transform = transform
        self.interpolation = cv2.INTER_LINEAR
        self.num_iterations = -1
        self.max_iterations = 30
        self.update_step_scale = 1.0
        self.failed = False
        self.eps = FLT_MAX
        self.max_eps = 0.2
        self.max_epse = 1e-4
        self.reference_image = None
        self.reference_mask = None
        self.current_image = None
        self.current_mask = None
        self.trace = None  # optional list; per-iteration records appended when not None

    def set_image_transform(self, t):
        self.transform = t

    def set_reference_image(self, image, mask):
        # ecc2.cc:585-609
        assert image.dtype == np.float32 and image.ndim == 2
        self.reference_image = image.copy()
        self.reference_mask = None if mask is None else mask.copy()
        return True

    def set_current_image(self, image, mask):
        # ecc2.cc:611-632
        assert image.dtype == np.float32 and image.ndim == 2
        self.current_image = image.copy()
        self.current_mask = None if mask is None else mask.copy()
        return True

    def release_current_image(self):
        self.current_image = None
        self.current_mask = None


class EccForwardAdditive(EccAlign):
    """c_ecc_forward_additive (ecc2.cc:1182-1365)."""

    def set_reference_image(self, image, mask):
        # ecc2.cc:1187-1211
        super().set_reference_image(image, mask)
        if self.reference_mask is not None:
            if cv2.countNonZero(self.reference_mask) == self.reference_mask.size:
                self.reference_mask = None
            else:
                self.reference_mask = cv2.erode(self.reference_mask, _SE5, borderType=cv2.BORDER_REPLICATE)
        return True

    def set_current_image(self, image, mask):
        # ecc2.cc:1214-1234
        super().set_current_image(image, mask)
        if self.current_mask is None:
            self.current_mask = np.full(self.current_image.shape, 255, dtype=np.uint8)
        elif cv2.countNonZero(self.current_mask) != self.current_mask.size:
            self.current_mask = cv2.erode(self.current_mask, _SE5, borderType=cv2.BORDER_REPLICATE)
        return True

    def align(self):
        # ecc2.cc:1247-1365
        self.failed = False
        self.num_iterations = -1
        if self.max_eps <= 0:
            self.max_eps = 1e-3
        t = self.transform
        f = self.reference_image
        size = _size(f)
        self.num_iterations = 0
        while True:
            cont = self.num_iterations < self.max_iterations
            self.num_iterations += 1
            if not cont:
                break
            rmap = t.create_remap(size)
            gw = cv2.remap(self.current_image, rmap, None, self.interpolation, borderMode=cv2.BORDER_REPLICATE)
            gxw, gyw = ecc_differentiate(gw)
            wmask = cv2.remap(self.current_mask, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
            wmask = cv2.compare(wmask, 255, cv2.CMP_GE)
            if self.reference_mask is not None:
                wmask = cv2.bitwise_and(wmask, self.reference_mask)
            iw = wmask == 0
            gxw[iw] = 0
            gyw[iw] = 0
            fMean, fStd = cv2.meanStdDev(f, mask=wmask)
            gMean, gStd = cv2.meanStdDev(gw, mask=wmask)
            stdev_ratio = float(gStd[0, 0]) / float(fStd[0, 0])
            jac = t.create_steepest_descent_images(gxw, gyw)
            H = ecc_compute_hessian_matrix(jac)
            ok, Hinv = cv2.invert(H, flags=cv2.DECOMP_CHOLESKY)
            if not ok:
                self.failed = True
                break
            # rhs = gw - r*f - (gMean - r*fMean)   (cv::scaleAdd then cv::subtract with a Scalar)
            rhs = cv2.scaleAdd(f, -stdev_ratio, gw)
            rhs = cv2.subtract(rhs, (float(gMean[0, 0]) - stdev_ratio * float(fMean[0, 0]), 0, 0, 0))
            rhs[iw] = 0
            ep = ecc_project_error_image(jac, rhs)
            dp = cv2.gemm(Hinv, ep, -self.update_step_scale, None, 0)   # -_update_step_scale * (H * ep)
            if self.trace is not None:
                self.trace.append(dict(H=H.copy(), ep=ep.copy(), dp=dp.copy(), p=t.parameters().copy(),
                                       r=stdev_ratio))
            t.set_parameters((t.parameters() + dp.reshape(-1)).astype(f32))
            self.eps = t.eps(dp, size)
            if self.eps < self.max_eps:
                break
        return not self.failed


class EccLM(EccAlign):
    """c_ecclm (ecc2.cc:1371-1650)."""

    def set_reference_image(self, image, mask):
        # ecc2.cc:1384-1406 (default erode border: BORDER_CONSTANT with +inf => border does not erode)
        super().set_reference_image(image, mask)
        if self.reference_mask is not None:
            self.reference_mask = cv2.erode(self.reference_mask, _SE5)
        return True

    def _compute_remap(self, params):
        # ecc2.cc:1444-1474
        size = _size(self.reference_image)
        remapped_image, remapped_mask = ecc_remap(self.transform, params, size, self.current_image,
                                                  self.current_mask, cv2.BORDER_REPLICATE)
        if self.reference_mask is not None:
            remapped_mask = cv2.bitwise_and(self.reference_mask, remapped_mask)
        rhs = cv2.subtract(remapped_image, self.reference_image)
        rhs[remapped_mask == 0] = 0
        self._remapped_image, self._remapped_mask, self._rhs = remapped_image, remapped_mask, rhs

    def _compute_rhs(self, params):
        # ecc2.cc:1476-1480
        self._compute_remap(params)
        self._rms = _norm_l2sqr(self._rhs)
        return self._rms

    def _compute_jac(self, params, recompute_remap):
        # ecc2.cc:1482-1525
        if recompute_remap:
            self._compute_remap(params)
            self._rms = _norm_l2sqr(self._rhs)
        gx, gy = ecc_differentiate(self._remapped_image, self._remapped_mask)
        J = self.transform.create_steepest_descent_images(gx, gy, params)
        v = ecc_project_error_image(J, self._rhs)
        H = ecc_compute_hessian_matrix(J)
        return self._rms, H, v

    def align(self):
        # ecc2.cc:1528-1650
        t = self.transform
        size = _size(self.reference_image)
        params = t.parameters().copy()
        M = params.size
        epsx, epse = self.max_eps, self.max_epse
        max_iterations = self.max_iterations
        lam = 0.1
        iteration = 0
        converged = False
        recompute_remap = True
        while iteration < max_iterations:
            err, H, v = self._compute_jac(params, recompute_remap)
            if err < 1:
                converged = True
                break
            Hp = H.copy()
            while True:
                cont = iteration < max_iterations
                iteration += 1
                if not cont:
                    break
                recompute_remap = True
                for i in range(M):
                    H[i, i] = f32((1 + lam) * float(Hp[i, i]))
                ok, deltap = cv2.solve(H, v, flags=cv2.DECOMP_CHOLESKY)
                if not ok:
                    # cv::solve returns false and leaves dst zero-filled for a non-SPD system
                    deltap = np.zeros((M, 1), dtype=f32)
                newparams = cv2.scaleAdd(deltap, -self.update_step_scale, params.reshape(-1, 1)).reshape(-1).astype(f32)
                self.eps = t.eps(deltap, size)
                if self.trace is not None:
                    self.trace.append(dict(H=Hp.copy(), v=v.copy(), dp=deltap.copy(), p=params.copy(), lam=lam, err=err))
                if self.eps <= epsx:
                    t.set_parameters(newparams)
                    params = t.parameters().copy()
                    converged = True
                    break
                newerr = self._compute_rhs(newparams)
                if newerr > err:
                    if lam > 1e6:
                        break
                    lam *= 10.0
                    continue
                t.set_parameters(newparams)
                params = t.parameters().copy()
                recompute_remap = False
                diff = err - newerr
                if diff < err * epse:
                    converged = True
                    break
                temp_d = cv2.gemm(Hp, deltap, -1, v, 2)
                dS = _dot(deltap, temp_d)
                rho = diff / abs(dS) if abs(dS) > float(f32(1e-9)) else diff
                if rho > 0.25:
                    lam = max(1e-8, 0.2 * lam)
                elif rho < 0.1:
                    lam = 1.0 if lam < 1.0 else lam * 10.0
                break
            if converged:
                break
        self.num_iterations = iteration
        return converged


class EccInverseCompositional(EccAlign):
    """c_ecc_inverse_compositional (ecc2.cc:1656-1788)."""

    def __init__(self, transform=None):
        super().__init__(transform)
        self._jac = None

    def set_image_transform(self, t):
        self._jac = None
        super().set_image_transform(t)

    def set_reference_image(self, image, mask):
        self._jac = None
        return super().set_reference_image(image, mask)

    def align(self):
        t = self.transform
        assert t.invertible()
        f = self.reference_image
        size = _size(f)
        params = t.parameters().copy()
        M = params.size
        RMA = f.size if self.reference_mask is None else cv2.countNonZero(self.reference_mask)
        if self._jac is None or len(self._jac) != M:
            gx, gy = ecc_differentiate(f, self.reference_mask)
            self._jac = t.create_steepest_descent_images(gx, gy)
            self._H = ecc_compute_hessian_matrix(self._jac)
        self.num_iterations = 0
        self.eps = FLT_MAX
        self.failed = False
        rmsold = FLT_MAX
        lam = self.update_step_scale
        while True:
            cont = self.num_iterations < self.max_iterations
            self.num_iterations += 1
            if not cont:
                break
            params = t.parameters().copy()
            remapped_image, remapped_mask = ecc_remap(t, params, size, self.current_image, self.current_mask)
            rhs = cv2.subtract(remapped_image, f)
            rhs[remapped_mask == 0] = 0
            CMA = cv2.countNonZero(remapped_mask)
            rmsnew = _ddiv(_norm_l2sqr(rhs) * (RMA * RMA), CMA * CMA)
            v = ecc_project_error_image(self._jac, rhs)
            ok, deltap = cv2.solve(self._H, (v * f32(_ddiv(RMA, CMA))).astype(f32), flags=cv2.DECOMP_CHOLESKY)
            if not ok:
                deltap = np.zeros((M, 1), dtype=f32)
            if rmsnew >= rmsold:
                break
            newparams = t.invert_and_compose(params, (f32(lam) * deltap).astype(f32))
            rmsold = rmsnew
            if self.trace is not None:
                self.trace.append(dict(v=v.copy(), dp=deltap.copy(), p=params.copy(), err=rmsnew))
            t.set_parameters(newparams)
            self.eps = t.eps(deltap, size)
            if self.eps < self.max_eps:
                break
        return True


class EccLMInverseCompositional(EccAlign):
    """c_ecclm_inverse_compositional (ecc2.cc:1794-2086)."""

    def __init__(self, transform=None):
        super().__init__(transform)
        self._jac = None
        self._reference_image_changed = True

    def set_reference_image(self, image, mask):
        self._reference_image_changed = True
        return super().set_reference_image(image, mask)

    def _ecc_remap(self, params, size):
        # ecc2.cc:1865-1892
        rmap = self.transform.create_remap(size, params)
        dst = cv2.remap(self.current_image, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        inv_dst_mask = cv2.remap(self._inv_current_mask, rmap, None, cv2.INTER_NEAREST,
                                 borderMode=cv2.BORDER_CONSTANT, borderValue=255)
        return dst, inv_dst_mask

    def _compute_rhs(self, params):
        # ecc2.cc:1894-1917
        f = self.reference_image
        size = _size(f)
        remapped, inv_mask = self._ecc_remap(params, size)
        if self._inv_reference_mask is not None:
            inv_mask = cv2.bitwise_or(self._inv_reference_mask, inv_mask)
        rhs = cv2.subtract(remapped, f)
        bad = cv2.countNonZero(inv_mask)
        self._CMA = f.size - bad
        rhs[inv_mask != 0] = 0
        self._rhs = rhs
        self._last_rms = _ddiv(_norm_l2sqr(rhs) * (self._RMA * self._RMA), self._CMA * self._CMA)
        return self._last_rms

    def _compute_v(self):
        # ecc2.cc:1919-1924
        v = ecc_project_error_image(self._jac, self._rhs)
        with np.errstate(invalid="ignore", over="ignore"):
            return (v * f32(_ddiv(self._RMA, self._CMA))).astype(f32)

    def align(self):
        # ecc2.cc:1926-2086
        t = self.transform
        assert t.invertible()
        f = self.reference_image
        size = _size(f)
        self._inv_reference_mask = None if self.reference_mask is None else cv2.bitwise_not(self.reference_mask)
        if self.current_mask is not None:
            self._inv_current_mask = cv2.bitwise_not(self.current_mask)
        else:
            self._inv_current_mask = np.zeros(self.current_image.shape, dtype=np.uint8)
        params = t.parameters().copy()
        M = params.size
        DBL_EPS = float(np.finfo(np.float64).eps)
        lam = 0.001
        dp = 0.0
        recompute_remap = True
        self._RMA = f.size if self.reference_mask is None else cv2.countNonZero(self.reference_mask)
        if self._jac is None or len(self._jac) != M or self._reference_image_changed:
            gx, gy = ecc_differentiate(f, self._inv_reference_mask, inverted=True)
            self._jac = t.create_steepest_descent_images(gx, gy)
            self._Hp = ecc_compute_hessian_matrix(self._jac)
            self._reference_image_changed = False
        Hp = self._Hp
        self.num_iterations = 0
        self.eps = FLT_MAX
        self.failed = False
        newerr = err = 0.0
        deltap = np.zeros((M, 1), dtype=f32)
        newparams = params
        while self.num_iterations < self.max_iterations:
            if recompute_remap:
                self._compute_rhs(params)
            v = self._compute_v()
            H = Hp.copy()
            err = self._last_rms
            while True:
                self.num_iterations += 1
                recompute_remap = True
                for i in range(M):
                    H[i, i] = f32((1 + lam) * float(Hp[i, i]))
                ok, deltap = cv2.solve(H, v, flags=cv2.DECOMP_CHOLESKY)
                if not ok:
                    deltap = np.zeros((M, 1), dtype=f32)
                newparams = t.invert_and_compose(params, (f32(self.update_step_scale) * deltap).astype(f32))
                newerr = self._compute_rhs(newparams)
                dp = t.eps(deltap, size)
                if self.trace is not None:
                    self.trace.append(dict(v=v.copy(), dp=deltap.copy(), p=params.copy(), lam=lam, err=err,
                                           newerr=newerr, eps=dp, cma=self._CMA, newp=np.asarray(newparams).copy()))
                if dp < self.max_eps:
                    break
                temp_d = cv2.gemm(Hp, deltap, -1, v, 2)
                dS = _dot(deltap, temp_d)
                rho = (err - newerr) / (dS if abs(dS) > DBL_EPS else 1)
                if rho > 0.25:
                    if lam > 1e-6:
                        lam = max(1e-6, lam / 5)
                elif rho > 0.1:
                    pass
                elif lam < 1:
                    lam = 1
                else:
                    lam *= 10
                if newerr < err:
                    break
                if not (self.num_iterations < self.max_iterations):
                    break
            if newerr < err:
                err = newerr
                recompute_remap = False
                t.set_parameters(newparams)
                params = t.parameters().copy()
            if dp < self.max_eps:
                break
        self.eps = dp
        return True


class EccH:
    """c_ecch coarse-to-fine driver (ecc2.cc:695-1176); options = c_ecch_options (ecc2.h:158-172)."""

    def __init__(self, transform=None, method=ECC_ALIGN_INVERSE_COMPOSITIONAL_LM, epsx=1e-5,
                 reference_smooth_sigma=1.0, input_smooth_sigma=1.0, update_step_scale=1.0,
                 interpolation=cv2.INTER_LINEAR, max_iterations=50, minimum_image_size=8, maxlevel=0):
        self.transform = transform
        self.method = method
        self.epsx = epsx
        self.reference_smooth_sigma = reference_smooth_sigma
        self.input_smooth_sigma = input_smooth_sigma
        self.update_step_scale = update_step_scale
        self.interpolation = interpolation
        self.max_iterations = max_iterations
        self.minimum_image_size = minimum_image_size
        self.maxlevel = maxlevel
        self.pyramid = []
        self.num_iterations = -1
        self.trace = None

    def set_image_transform(self, t):
        # ecc2.cc:723-732
        if self.transform is not t:
            self.transform = t
            for m in self.pyramid:
                m.set_image_transform(t)

    def _create_ecc_align(self, epsx):
        # ecc2.cc:944-970
        cls = {ECC_ALIGN_FORWARD_ADDITIVE: EccForwardAdditive,
               ECC_ALIGN_INVERSE_COMPOSITIONAL: EccInverseCompositional,
               ECC_ALIGN_INVERSE_COMPOSITIONAL_LM: EccLMInverseCompositional}.get(self.method, EccLM)
        e = cls()
        e.set_image_transform(self.transform)
        e.interpolation = self.interpolation
        e.max_iterations = self.max_iterations
        e.update_step_scale = self.update_step_scale
        e.max_eps = epsx
        return e

    @staticmethod
    def _smooth(image, sigma):
        # ecc2.cc:998-1006
        if sigma > 0:
            ksize = max(3, 2 * int(3 * sigma) + 1)
            G = cv2.getGaussianKernel(ksize, sigma)
            image = cv2.sepFilter2D(image, -1, G, G, borderType=cv2.BORDER_REPLICATE)
        return image

    @staticmethod
    def _downscale(image, mask, next_size):
        # ecc2.cc:972-982
        image = cv2.pyrDown(image, dstsize=next_size)
        if mask is not None:
            mask = cv2.resize(mask, next_size, interpolation=cv2.INTER_NEAREST)
        return image, mask

    def set_reference_image(self, reference_image, reference_mask):
        # ecc2.cc:984-1057
        self.num_iterations = -1
        image, mask = ecc_convert_input_image(reference_image, reference_mask)
        image = self._smooth(image, self.reference_smooth_sigma)
        min_image_size = max(4, self.minimum_image_size)
        test_size = _size(image)
        required_lvls = 1
        while True:
            if self.maxlevel >= 0 and required_lvls >= max(0, self.maxlevel):
                break
            nxt = compute_next_pyramid_layer_size(test_size)
            if nxt[0] < min_image_size or nxt[1] < min_image_size:
                break
            test_size = nxt
            required_lvls += 1
        if len(self.pyramid) != required_lvls:
            self.pyramid = []
            epsx = self.epsx
            for _ in range(required_lvls):
                self.pyramid.append(self._create_ecc_align(epsx))
                epsx *= 2
        self.pyramid[0].set_reference_image(image, mask)
        for lvl in range(1, required_lvls):
            nxt = compute_next_pyramid_layer_size(_size(image))
            image, mask = self._downscale(image, mask, nxt)
            self.pyramid[lvl].set_reference_image(image, mask)
        return True

    def set_current_image(self, current_image, current_mask):
        # ecc2.cc:1060-1120
        assert self.pyramid
        self.num_iterations = -1
        image, mask = ecc_convert_input_image(current_image, current_mask)
        image = self._smooth(image, self.input_smooth_sigma)
        lvls = len(self.pyramid)
        lvl = 0
        while lvl < lvls:
            self.pyramid[lvl].set_current_image(image, mask)
            if lvl < lvls - 1:
                nxt = compute_next_pyramid_layer_size(_size(image))
                if nxt[0] < 4 or nxt[1] < 4:
                    break
                image, mask = self._downscale(image, mask, nxt)
            lvl += 1
        while lvl < lvls:
            self.pyramid[lvl].release_current_image()
            lvl += 1
        return True

    def align(self, current_image=None, current_mask=None):
        # ecc2.cc:1122-1176
        if current_image is not None:
            self.set_current_image(current_image, current_mask)
        lvls = len(self.pyramid)
        lvl = lvls - 1
        while lvl > 0 and self.pyramid[lvl].current_image is None:
            lvl -= 1
        if lvl > 0:
            w0 = self.pyramid[0].reference_image.shape[1]
            w1 = self.pyramid[lvl].reference_image.shape[1]
            self.transform.scale_transfrom(float(w1) / float(w0))
        self.num_iterations = 0
        while lvl >= 0:
            e = self.pyramid[lvl]
            e.trace = [] if self.trace is not None else None
            ok = e.align()
            if self.trace is not None:
                self.trace.append((lvl, e.trace))
            if ok and lvl > 0:
                w0 = self.pyramid[lvl].reference_image.shape[1]
                w1 = self.pyramid[lvl - 1].reference_image.shape[1]
                self.transform.scale_transfrom(float(w1) / float(w0))
            self.num_iterations += e.num_iterations
            lvl -= 1
        return True

    def eps(self):
        return -1 if not self.pyramid else self.pyramid[0].eps

    def reference_image(self):
        return self.pyramid[0].reference_image

    def reference_mask(self):
        return self.pyramid[0].reference_mask

    def current_image(self):
        return self.pyramid[0].current_image

    def current_mask(self):
        return self.pyramid[0].current_mask

    def create_remap(self):
        # ecc2.cc:921-942
        return self.transform.create_remap(_size(self.reference_image()))
