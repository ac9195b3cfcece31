"""
Oracle restatement of c_eccflow, the dense smooth optical flow that refines the registration map per pixel.

Follows /root/reference/core/proc/image_registration/:
  c_eccflow_options                    ecc2.h:515-527
  c_eccflow::convert_input_images      ecc2.cc:2220-2235
  c_eccflow::compute_uv (one level)    ecc2.cc:2237-2398
  c_eccflow::avgdown / avgp            ecc2.cc:2400-2430
  c_eccflow::downscale / upscale       ecc2.cc:2432-2480
  c_eccflow::set_reference_image       ecc2.cc:2494-2672
  c_eccflow::setup_input_image         ecc2.cc:2674-2768
  c_eccflow::compute_uv / compute      ecc2.cc:2773-2865
  ecc_remap_to_optflow / ecc_flow_to_remap   ecc2.cc:402-500
  c_eccflow_registration_options       c_frame_registration.h:88-100 (the values c_frame_registration passes,
                                       c_frame_registration.cc:637-660)

Every OpenCV primitive the reference calls is called here through cv2 with the same arguments.
Test infrastructure only (see oracle/__init__.py).  PARITY UNPINNED: the reference holds no vectors for this class.
"""
import numpy as np
import cv2

from .ecc import ecc_differentiate

f32 = np.float32

# ecc2.h: enum ECCFlowDownscaleMethod
DOWNSCALE_RECURSIVE_RESIZE = 0
DOWNSCALE_FULL_RESIZE = 1
DOWNSCALE_PYRAMID = 2

_G3 = cv2.getGaussianKernel(3, 0, cv2.CV_32F)


class EccFlowOptions:
    """c_eccflow_options defaults (ecc2.h:515-527)."""

    def __init__(self, **kw):
        self.input_smooth_sigma = 0.0       # unused by the current reference code
        self.reference_smooth_sigma = 0.0   # unused by the current reference code
        self.update_multiplier = 1.5
        self.scale_factor = 0.5
        self.noise_level = -1.0
        self.max_iterations = 1
        self.support_scale = 5
        self.min_image_size = 4
        self.max_pyramid_level = -1
        self.downscale = DOWNSCALE_RECURSIVE_RESIZE
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def registration_options(**kw):
    """The values c_frame_registration::setup_reference_frame hands to c_eccflow (c_frame_registration.h:88-100)."""
    d = dict(update_multiplier=1.5, input_smooth_sigma=0.0, reference_smooth_sigma=0.0, noise_level=-1.0, scale_factor=0.75,
             max_iterations=3, support_scale=4, min_image_size=-1, max_pyramid_level=-1, downscale=DOWNSCALE_RECURSIVE_RESIZE)
    d.update(kw)
    return EccFlowOptions(**d)


def remap_to_optflow(rmap):
    """ecc2.cc:402-447"""
    h, w = rmap.shape[:2]
    flow = np.empty_like(rmap, dtype=f32)
    flow[..., 0] = rmap[..., 0] - np.arange(w, dtype=f32)[None, :]
    flow[..., 1] = rmap[..., 1] - np.arange(h, dtype=f32)[:, None]
    return flow


def flow_to_remap(flow):
    """ecc2.cc:452-500"""
    h, w = flow.shape[:2]
    rmap = np.empty_like(flow, dtype=f32)
    rmap[..., 0] = flow[..., 0] + np.arange(w, dtype=f32)[None, :]
    rmap[..., 1] = flow[..., 1] + np.arange(h, dtype=f32)[:, None]
    return rmap


def _size(img):
    return (img.shape[1], img.shape[0])


class _Entry:
    __slots__ = ("current_image", "reference_image", "current_mask", "reference_mask", "Ix", "Iy", "D")

    def __init__(self):
        self.current_image = self.reference_image = self.current_mask = self.reference_mask = None
        self.Ix = self.Iy = self.D = None


class EccFlow:
    """c_eccflow (ecc2.h:548-662)."""

    def __init__(self, opts: EccFlowOptions = None):
        self.opts = opts if opts is not None else EccFlowOptions()
        self.pyramid = []
        self.uv = None

    # ---- helpers -----------------------------------------------------------------------------------------
    def avgdown_size(self, size):
        w, h = size
        for _ in range(self.opts.support_scale):
            w, h = (w + 1) // 2, (h + 1) // 2
        return (w, h)

    def avgdown(self, src):
        """ecc2.cc:2400-2418"""
        dst = cv2.resize(src, self.avgdown_size(_size(src)), interpolation=cv2.INTER_AREA)
        if dst.ndim == 2 and src.ndim == 3:
            dst = dst.reshape(dst.shape[0], dst.shape[1], src.shape[2])
        return cv2.sepFilter2D(dst, -1, _G3, _G3, borderType=cv2.BORDER_REPLICATE)

    def avgp(self, a, b):
        """ecc2.cc:2425-2430"""
        return self.avgdown(cv2.multiply(a, b))

    def downscale(self, src, src_mask, dst_size):
        """ecc2.cc:2432-2458; dst_size = (w, h)"""
        if self.opts.downscale == DOWNSCALE_PYRAMID:
            dst = cv2.pyrDown(src, dstsize=dst_size)
        else:
            dst = cv2.resize(src, dst_size, interpolation=cv2.INTER_AREA)
        dst_mask = None
        if src_mask is not None:
            dst_mask = cv2.resize(src_mask, dst_size, interpolation=cv2.INTER_NEAREST)
        return dst, dst_mask

    @staticmethod
    def upscale(src, dst_size):
        """ecc2.cc:2460-2480 (the flow has no mask)"""
        return cv2.resize(src, dst_size, interpolation=cv2.INTER_CUBIC)

    # ---- reference side ----------------------------------------------------------------------------------
    def level_sizes(self, image_size):
        """The size recursion of set_reference_image (ecc2.cc:2520-2672) -> list of ((w, h), source level) where source
        level is the level the image is reduced from (0 = pyramid front for the big-aspect-ratio / full-resize rule)."""
        o = self.opts
        min_image_size = max(4, o.min_image_size)
        w0, h0 = image_size
        big_aspect_ratio = (max(w0, h0) // min(w0, h0)) >= 2
        sizes = [((w0, h0), -1)]
        lvl = 0
        while True:
            if o.max_pyramid_level >= 0 and lvl >= o.max_pyramid_level:
                break
            lvl += 1
            pw, ph = sizes[-1][0]
            if o.downscale in (DOWNSCALE_RECURSIVE_RESIZE, DOWNSCALE_FULL_RESIZE):
                nxt = (max(o.min_image_size, int((pw + 1) * o.scale_factor)), max(o.min_image_size, int((ph + 1) * o.scale_factor)))
                if nxt == (pw, ph) or max(nxt) <= min_image_size:
                    break
                if o.downscale == DOWNSCALE_FULL_RESIZE:
                    src = 0
                else:
                    src = 0 if (big_aspect_ratio and min(nxt) <= min_image_size + 1) else lvl - 1
            else:
                nxt = (max(o.min_image_size, (pw + 1) // 2), max(o.min_image_size, (ph + 1) // 2))
                if nxt == (pw, ph) or min(nxt) <= min_image_size:
                    break
                src = lvl - 1
            sizes.append((nxt, src))
        return sizes

    def set_reference_image(self, reference_image, reference_mask=None):
        o = self.opts
        assert reference_image.ndim == 2
        noise_level = o.noise_level if o.noise_level >= 0 else 1e-3
        self.pyramid = []
        for lvl, (size, src) in enumerate(self.level_sizes(_size(reference_image))):
            e = _Entry()
            if lvl == 0:
                e.reference_image = np.ascontiguousarray(reference_image, dtype=f32)
                e.reference_mask = None if reference_mask is None else reference_mask.copy()
            else:
                s = self.pyramid[src]
                e.reference_image, e.reference_mask = self.downscale(s.reference_image, s.reference_mask, size)
            e.Ix, e.Iy = ecc_differentiate(e.reference_image)
            Ixx, Ixy, Iyy = self.avgp(e.Ix, e.Ix), self.avgp(e.Ix, e.Iy), self.avgp(e.Iy, e.Iy)
            # "FIXME: this regularization term estimation looks crazy" (ecc2.cc:2613): float(pow(double, 4))
            reg = f32((1e-5 * noise_level / (1 << lvl)) ** 4) if noise_level > 0 else f32(0)
            um = f32(o.update_multiplier)
            det = np.abs((Ixx * Iyy).astype(f32) - (Ixy * Ixy).astype(f32)).astype(f32)
            idet = (um / (det + reg).astype(f32)).astype(f32)
            e.D = np.stack([Ixx, Ixy, Iyy, idet], axis=-1).astype(f32)
            self.pyramid.append(e)
        return True

    # ---- current side ------------------------------------------------------------------------------------
    def setup_input_image(self, input_image, input_mask=None):
        assert self.pyramid, "set_reference_image() must be called first"
        sizes = self.level_sizes(_size(self.pyramid[0].reference_image))
        for lvl, e in enumerate(self.pyramid):
            if lvl == 0:
                e.current_image = np.ascontiguousarray(input_image, dtype=f32)
                e.current_mask = None if input_mask is None else input_mask.copy()
            else:
                s = self.pyramid[sizes[lvl][1]]
                e.current_image, e.current_mask = self.downscale(s.current_image, s.current_mask, _size(e.reference_image))
        return True

    def _compute_uv_level(self, e, rmap):
        """ecc2.cc:2237-2398"""
        W = cv2.remap(e.current_image, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)
        M = None
        if e.current_mask is not None:
            M = cv2.remap(e.current_mask, rmap, None, cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT)
        if e.reference_mask is not None:
            M = e.reference_mask if M is None else cv2.bitwise_and(e.reference_mask, M)
        It = (e.reference_image - W).astype(f32)
        Itxy = np.stack([(It * e.Ix).astype(f32), (It * e.Iy).astype(f32)], axis=-1)
        if M is not None:
            Itxy[M == 0] = 0
        Itxy = self.avgdown(Itxy)
        a00, a01, a11, det = e.D[..., 0], e.D[..., 1], e.D[..., 2], e.D[..., 3]
        b0, b1 = Itxy[..., 0], Itxy[..., 1]
        u = (det * ((a11 * b0).astype(f32) - (a01 * b1).astype(f32)).astype(f32)).astype(f32)
        v = (det * ((a00 * b1).astype(f32) - (a01 * b0).astype(f32)).astype(f32)).astype(f32)
        cuv = np.stack([u, v], axis=-1)
        return cv2.resize(cuv, _size(W), interpolation=cv2.INTER_CUBIC)

    def compute_uv(self, input_image, rmap, input_mask=None):
        """ecc2.cc:2773-2836; rmap: H x W x 2 float32 or None."""
        self.setup_input_image(input_image, input_mask)
        first, last = self.pyramid[0], self.pyramid[-1]
        fsz, lsz = _size(first.reference_image), _size(last.reference_image)
        if rmap is None:
            uv = np.zeros((lsz[1], lsz[0], 2), f32)
        else:
            assert _size(rmap) == fsz
            uv = remap_to_optflow(rmap)
            uv = cv2.resize(uv, lsz, interpolation=cv2.INTER_CUBIC)
            # cv::multiply(Mat2f, Scalar): the scalar is narrowed to float
            uv = (uv * np.array([f32(lsz[0] / fsz[0]), f32(lsz[1] / fsz[1])], f32)).astype(f32)
        n = len(self.pyramid)
        for i in range(n - 1, -1, -1):
            e = self.pyramid[i]
            if i < n - 1:
                csz, psz = _size(e.current_image), _size(self.pyramid[i + 1].current_image)
                uv = (uv * np.array([f32(csz[0] / psz[0]), f32(csz[1] / psz[1])], f32)).astype(f32)
                uv = self.upscale(uv, csz)
            for _ in range(self.opts.max_iterations):
                cuv = self._compute_uv_level(e, flow_to_remap(uv))
                uv = (uv + cuv).astype(f32)
        self.uv = uv
        return True

    def compute(self, input_image, rmap, input_mask=None):
        """ecc2.cc:2855-2865 -> the refined map."""
        self.compute_uv(input_image, rmap, input_mask)
        return flow_to_remap(self.uv)
