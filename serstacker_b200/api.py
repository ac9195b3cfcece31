"""
Python harness over the C ABI that mirrors the reference's operator surface for the stacking hot path
(same class / method names and argument meaning, see the citations), so the parity tests read like code
written against the reference.  The product host layer is the C++ adapter (serstacker_b200/host/ssk_adapter.h);
this module only marshals numpy arrays into ssk_mat and calls libssk.so.  No computation happens here.
"""
import ctypes as C
import os
import numpy as np

from . import capi
from .capi import (check, mat, ref, ssk_transform, ssk_ecc_status, ssk_mat, SskError)

f32 = np.float32


class c_image_transform:
    """c_image_transform by value (core/proc/image_registration/c_image_transform.h:42-117)."""

    def __init__(self, motion_type):
        self.t = ssk_transform()
        check(capi.lib.ssk_transform_init(C.byref(self.t), motion_type))

    @property
    def motion_type(self):
        return self.t.motion_type

    def parameters(self):
        return self.t.parameters()

    def set_parameters(self, p):
        p = np.asarray(p, dtype=f32).reshape(-1)
        assert p.size == self.t.nparams
        for i, v in enumerate(p):
            self.t.params[i] = float(v)
        return True

    def scale_transfrom(self, factor):
        check(capi.lib.ssk_transform_scale(C.byref(self.t), float(factor)))

    def eps(self, dp, size):
        """c_image_transform::eps(dp, image_size), size = (width, height)."""
        dp = np.ascontiguousarray(dp, dtype=f32).reshape(-1)
        out = C.c_double()
        check(capi.lib.ssk_transform_eps(C.byref(self.t), dp.ctypes.data_as(C.POINTER(C.c_float)), dp.size, int(size[1]), int(size[0]),
                                         C.byref(out)))
        return out.value

    def invert_and_compose(self, dp):
        """c_image_transform::invert_and_compose(parameters(), dp) -> new parameter vector."""
        dp = np.ascontiguousarray(dp, dtype=f32).reshape(-1)
        out = np.empty(self.t.nparams, dtype=f32)
        check(capi.lib.ssk_transform_invert_and_compose(C.byref(self.t), dp.ctypes.data_as(C.POINTER(C.c_float)), dp.size,
                                                        out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def remap_points(self, rpts):
        """c_image_transform::remap(parameters(), rpts, cpts): (n, 2) reference points -> current-frame points."""
        rpts = np.ascontiguousarray(rpts, dtype=f32).reshape(-1, 2)
        cpts = np.empty_like(rpts)
        check(capi.lib.ssk_transform_remap_points(C.byref(self.t), rpts.ctypes.data_as(C.POINTER(C.c_float)), rpts.shape[0],
                                                  cpts.ctypes.data_as(C.POINTER(C.c_float))))
        return cpts

    def create_remap(self, size):
        w, h = size
        rmap = np.empty((h, w, 2), dtype=f32)
        m = mat(rmap)
        check(capi.lib.ssk_transform_create_remap(C.byref(self.t), h, w, C.byref(m)))
        return rmap


def create_image_transform(motion_type):
    """image_transform.cc:34-63."""
    return c_image_transform(motion_type)


def remap(transform, rmap, src, want_mask=False, src_mask=None, interpolation=capi.INTER_LINEAR,
          border_mode=capi.BORDER_REFLECT101, border_value=(0, 0, 0, 0), dst=None, size=None):
    """c_frame_registration::base_remap (c_frame_registration.cc:1265-1386) -> (dst, dst_mask)."""
    if size is None:
        size = rmap.shape[:2] if rmap is not None else src.shape[:2]
    out = None
    if src is not None:
        out = dst if dst is not None else np.zeros(tuple(size) + src.shape[2:], dtype=f32)
    omask = np.zeros(tuple(size), dtype=np.uint8) if want_mask else None
    ms, mo, mm, mk, mr = mat(src), mat(out), mat(src_mask), mat(omask), mat(rmap)
    bv = (C.c_double * 4)(*[float(v) for v in border_value])
    check(capi.lib.ssk_remap(None if transform is None else C.byref(transform.t), ref(mr), ref(ms), ref(mo), ref(mm),
                             ref(mk), interpolation, border_mode, bv))
    return out, omask


class c_ecch:
    """c_ecch (ecc2.h:193-308)."""

    def __init__(self, transform=None, method=capi.ECC_INVERSE_COMPOSITIONAL_LM, **opts):
        o = capi.ssk_ecch_options()
        capi.lib.ssk_ecch_options_default(C.byref(o))
        o.method = method
        for k, v in opts.items():
            assert hasattr(o, k), k
            setattr(o, k, v)
        self._h = C.c_void_p()
        check(capi.lib.ssk_ecch_create(C.byref(o), C.byref(self._h)))
        self.transform = transform
        self.status = ssk_ecc_status()

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_ecch_destroy(self._h)
            self._h = None

    def set_image_transform(self, t):
        self.transform = t

    def set_reference_image(self, image, mask=None):
        m = mat(np.ascontiguousarray(image))
        check(capi.lib.ssk_ecch_set_reference_image(self._h, C.byref(m), ref(mat(mask))))
        return True

    def align(self, image, mask=None):
        m = mat(np.ascontiguousarray(image))
        check(capi.lib.ssk_ecch_align(self._h, C.byref(m), ref(mat(mask)), C.byref(self.transform.t), C.byref(self.status)))
        return True

    def eps(self):
        return self.status.eps

    def num_iterations(self):
        return self.status.num_iterations

    def num_levels(self):
        return capi.lib.ssk_ecch_num_levels(self._h)

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(capi.lib.ssk_ecch_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def _image(self, which, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), dtype=f32)
        m = mat(out)
        check(capi.lib.ssk_ecch_get_image(self._h, which, level, C.byref(m)))
        return out

    def reference_image(self, level=0):
        return self._image(0, level)

    def current_image(self, level=0):
        return self._image(1, level)


def registration_options(**kw):
    """c_image_registration_options with the reference defaults (c_frame_registration.h:47-64, 119-136);
    keyword `ecc` is a dict of c_ecc_registration_options fields, `eccflow` one of c_eccflow_registration_options fields."""
    o = capi.ssk_registration_options()
    capi.lib.ssk_registration_options_default(C.byref(o))
    o.enable_ecc_registration = 1
    ecc = kw.pop("ecc", {})
    for k, v in kw.pop("eccflow", {}).items():
        assert hasattr(o.eccflow, k), k
        setattr(o.eccflow, k, v)
    for k, v in kw.items():
        if k == "border_value":
            for i in range(4):
                o.border_value[i] = float(v[i]) if i < len(v) else 0.0
        else:
            assert hasattr(o, k), k
            setattr(o, k, v)
    for k, v in ecc.items():
        assert hasattr(o.ecc, k), k
        setattr(o.ecc, k, v)
    return o


def _upscaled_shape(option, shape):
    uc, ur = C.c_int(), C.c_int()
    check(capi.lib.ssk_upscale_size(option, shape[1], shape[0], C.byref(uc), C.byref(ur)))
    return (ur.value, uc.value) + tuple(shape[2:])


def upscale_image(option, src, srcmask=None):
    """c_image_stacking_pipeline::upscale_image (c_image_stacking_pipeline.cc:1949-2000) -> (dst, dstmask)."""
    dst = None if src is None else np.empty(_upscaled_shape(option, src.shape), dtype=f32)
    dmask = None if srcmask is None else np.empty(_upscaled_shape(option, srcmask.shape), dtype=np.uint8)
    check(capi.lib.ssk_upscale_image(option, ref(mat(None if src is None else np.ascontiguousarray(src))), ref(mat(srcmask)), ref(mat(dst)), ref(mat(dmask))))
    return dst, dmask


def upscale_remap(option, srcmap):
    """c_image_stacking_pipeline::upscale_remap (c_image_stacking_pipeline.cc:1869-1905)."""
    dst = np.empty(_upscaled_shape(option, srcmap.shape), dtype=f32)
    check(capi.lib.ssk_upscale_remap(option, ref(mat(np.ascontiguousarray(srcmap))), ref(mat(dst))))
    return dst


def upscale_optflow(option, srcmap):
    """c_image_stacking_pipeline::upscale_optflow (c_image_stacking_pipeline.cc:1907-1946)."""
    dst = np.empty(_upscaled_shape(option, srcmap.shape), dtype=f32)
    check(capi.lib.ssk_upscale_optflow(option, ref(mat(np.ascontiguousarray(srcmap))), ref(mat(dst))))
    return dst


class c_canvas_average:
    """c_canvas_average (core/average/c_frame_accumulation.h:65-137)."""

    def __init__(self, interpolation=capi.INTER_LINEAR, canvas_size=None):
        self._h = C.c_void_p()
        check(capi.lib.ssk_canvas_create(interpolation, C.byref(self._h)))
        if canvas_size is not None:
            self.setCanvasSize(canvas_size)

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_canvas_destroy(self._h)
            self._h = None

    def setCanvasSize(self, size):
        check(capi.lib.ssk_canvas_set_canvas_size(self._h, int(size[0]), int(size[1])))

    def add(self, current_image, current_weights_or_mask=None, rmap=None, new_canvas_bbox=None):
        bb = None if new_canvas_bbox is None else (C.c_int * 4)(*[int(v) for v in new_canvas_bbox])
        rc = capi.lib.ssk_canvas_add(self._h, ref(mat(np.ascontiguousarray(current_image))), ref(mat(current_weights_or_mask)),
                                     ref(mat(rmap)), bb)
        if rc == capi.SSK_ERR_INVALID and "ROI is empty" in capi.last_error():
            return False                    # c_frame_accumulation.cc:378-381
        check(rc)
        return True

    def accumulator_size(self):
        v = [C.c_int() for _ in range(3)]
        check(capi.lib.ssk_canvas_size(self._h, *[C.byref(x) for x in v]))
        return v[0].value, v[1].value, v[2].value

    def last_bbox(self):
        bb = (C.c_int * 4)()
        check(capi.lib.ssk_canvas_last_bbox(self._h, bb))
        return tuple(bb)

    def accumulated_frames(self):
        return capi.lib.ssk_canvas_accumulated_frames(self._h)

    def compute(self, dscale=1.0, rbbox=None):
        w, h, cn = self.accumulator_size()
        if rbbox is not None and rbbox[2] > 0 and rbbox[3] > 0:
            x0, y0 = max(rbbox[0], 0), max(rbbox[1], 0)
            x1, y1 = min(rbbox[0] + rbbox[2], w), min(rbbox[1] + rbbox[3], h)
            w, h = x1 - x0, y1 - y0
        avg = np.empty((h, w) if cn == 1 else (h, w, cn), dtype=f32)
        mask = np.empty((h, w), dtype=np.uint8)
        bb = None if rbbox is None else (C.c_int * 4)(*[int(v) for v in rbbox])
        ma, mm = mat(avg), mat(mask)
        check(capi.lib.ssk_canvas_compute(self._h, C.byref(ma), C.byref(mm), float(dscale), bb))
        return avg, mask

    def clear(self):
        check(capi.lib.ssk_canvas_clear(self._h))


def eccflow_options(registration_defaults=False, **kw):
    """c_eccflow_options (ecc2.h:515-527) or, with registration_defaults, the values c_frame_registration hands to c_eccflow
    (c_eccflow_registration_options, c_frame_registration.h:88-100)."""
    o = capi.ssk_eccflow_options()
    (capi.lib.ssk_eccflow_registration_options_default if registration_defaults else capi.lib.ssk_eccflow_options_default)(C.byref(o))
    for k, v in kw.items():
        assert hasattr(o, k), k
        setattr(o, k, v)
    return o


class c_eccflow:
    """c_eccflow (ecc2.h:548-662)."""

    def __init__(self, options=None):
        self.options = options if options is not None else eccflow_options()
        self._h = C.c_void_p()
        check(capi.lib.ssk_eccflow_create(C.byref(self.options), C.byref(self._h)))
        self._shape = None

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_eccflow_destroy(self._h)
            self._h = None

    def set_reference_image(self, reference_image, reference_mask=None):
        m = mat(np.ascontiguousarray(reference_image))
        check(capi.lib.ssk_eccflow_set_reference_image(self._h, C.byref(m), ref(mat(reference_mask))))
        self._shape = reference_image.shape[:2]
        return True

    def compute(self, input_image, rmap=None, input_mask=None):
        """c_eccflow::compute(input_image, rmap, input_mask): returns the refined map (rmap = None: empty initial map)."""
        out = np.zeros(self._shape + (2,), dtype=f32) if rmap is None else np.ascontiguousarray(rmap, dtype=f32).copy()
        m, mo = mat(np.ascontiguousarray(input_image)), mat(out)
        check(capi.lib.ssk_eccflow_compute(self._h, C.byref(m), ref(mat(input_mask)), C.byref(mo), 0 if rmap is None else 1))
        return out

    def current_uv(self):
        out = np.empty(self._shape + (2,), dtype=f32)
        mo = mat(out)
        check(capi.lib.ssk_eccflow_get_uv(self._h, C.byref(mo)))
        return out

    def num_levels(self):
        return capi.lib.ssk_eccflow_num_levels(self._h)

    def level_size(self, level):
        v = [C.c_int() for _ in range(4)]
        check(capi.lib.ssk_eccflow_level_size(self._h, level, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def pyramid_image(self, which, level):
        """which: 0 reference_image, 1 current_image, 2 Ix, 3 Iy, 4 D (current_pyramid(), ecc2.h:555-560)."""
        w, h, gw, gh = self.level_size(level)
        out = np.empty((gh, gw, 4), dtype=f32) if which == 4 else np.empty((h, w), dtype=f32)
        mo = mat(out)
        check(capi.lib.ssk_eccflow_get_image(self._h, which, level, C.byref(mo)))
        return out


class c_frame_registration:
    """c_frame_registration, ECC branch (c_frame_registration.h:210-354)."""

    def __init__(self, options):
        self.options = options
        self._h = C.c_void_p()
        check(capi.lib.ssk_reg_create(C.byref(options), C.byref(self._h)))
        self.status = ssk_ecc_status()
        self.transform = ssk_transform()
        self._ref_shape = None

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_reg_destroy(self._h)
            self._h = None

    def setup_reference_frame(self, image, mask=None, bpp=0):
        m = mat(np.ascontiguousarray(image))
        check(capi.lib.ssk_reg_setup_reference_frame(self._h, C.byref(m), ref(mat(mask)), bpp))
        self._ref_shape = image.shape[:2]
        return True

    def register_frame(self, image, mask=None, bpp=0):
        """Returns True/False like the reference (False: low correlation / solver failure)."""
        m = mat(np.ascontiguousarray(image))
        rc = capi.lib.ssk_reg_register_frame(self._h, C.byref(m), ref(mat(mask)), bpp, C.byref(self.transform), C.byref(self.status))
        if rc == capi.SSK_ERR_NOT_REGISTERED:
            return False
        check(rc)
        return True

    def image_transform_parameters(self):
        return self.transform.parameters()

    def current_remap(self):
        h, w = self._ref_shape
        rmap = np.empty((h, w, 2), dtype=f32)
        m = mat(rmap)
        check(capi.lib.ssk_reg_get_current_remap(self._h, C.byref(m)))
        return rmap

    def custom_remap(self, rmap, src, src_mask=None, want_mask=True, interpolation=-1, border_mode=-1,
                     border_value=(0, 0, 0, 0)):
        size = rmap.shape[:2] if rmap is not None else self._ref_shape
        out = np.zeros(tuple(size) + src.shape[2:], dtype=f32) if src is not None else None
        omask = np.zeros(tuple(size), dtype=np.uint8) if want_mask else None
        bv = (C.c_double * 4)(*[float(v) for v in border_value])
        check(capi.lib.ssk_reg_remap(self._h, ref(mat(rmap)), ref(mat(src)), ref(mat(out)), ref(mat(src_mask)), ref(mat(omask)),
                                     interpolation, border_mode, bv))
        return out, omask

    def remap(self, src, src_mask=None, **kw):
        return self.custom_remap(None, src, src_mask, **kw)


class c_frame_accumulation:
    """c_frame_accumulation (core/average/c_frame_accumulation.h:14-37)."""
    kind = capi.ACC_WEIGHTED_AVERAGE

    def __init__(self, handle=None):
        self._own = handle is None
        self._h = C.c_void_p()
        if handle is None:
            check(capi.lib.ssk_acc_create(self.kind, C.byref(self._h)))
        else:
            self._h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_acc_destroy(self._h)
            self._h = None

    def add(self, src, weights=None, bpp=0):
        check(capi.lib.ssk_acc_add(self._h, C.byref(mat(np.ascontiguousarray(src))),
                                   ref(mat(None if weights is None else np.ascontiguousarray(weights))), bpp))
        return True

    def accumulator_size(self):
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        check(capi.lib.ssk_acc_size(self._h, C.byref(w), C.byref(h), C.byref(c)))
        return w.value, h.value, c.value

    def accumulated_frames(self):
        return capi.lib.ssk_acc_frames(self._h)

    def compute(self, dscale=1.0, want_mask=True):
        w, h, c = self.accumulator_size()
        avg = np.empty((h, w) if c == 1 else (h, w, c), dtype=f32)
        mask = np.empty((h, w), dtype=np.uint8) if want_mask else None
        check(capi.lib.ssk_acc_compute(self._h, C.byref(mat(avg)), ref(mat(mask)), float(dscale)))
        return avg, mask

    def compute_inpainted(self, dscale=1.0, max_levels=100):
        """compute() + average_pyramid_inpaint on the device (c_image_stacking_pipeline.cc:742-767)."""
        w, h, c = self.accumulator_size()
        avg = np.empty((h, w) if c == 1 else (h, w, c), dtype=f32)
        mask = np.empty((h, w), dtype=np.uint8)
        check(capi.lib.ssk_acc_compute_inpainted(self._h, C.byref(mat(avg)), C.byref(mat(mask)), float(dscale), int(max_levels)))
        return avg, mask

    def get_acc_counters(self):
        w, h, c = self.accumulator_size()
        wc = 3 if self.kind == capi.ACC_BAYER_AVERAGE else 1
        out = np.empty((h, w) if wc == 1 else (h, w, wc), dtype=f32)
        check(capi.lib.ssk_acc_get_counters(self._h, C.byref(mat(out))))
        return out

    def reinitialize(self, src, accw):
        check(capi.lib.ssk_acc_reinitialize(self._h, C.byref(mat(np.ascontiguousarray(src, dtype=f32))),
                                            C.byref(mat(np.ascontiguousarray(accw, dtype=f32)))))
        return True

    def clear(self):
        check(capi.lib.ssk_acc_clear(self._h))


class c_weigthed_average(c_frame_accumulation):
    """c_weigthed_average (c_frame_accumulation.h:39-63) [sic: the reference's spelling]."""
    kind = capi.ACC_WEIGHTED_AVERAGE


class c_bayer_average(c_frame_accumulation):
    """c_bayer_average (c_frame_accumulation.h:222-262)."""
    kind = capi.ACC_BAYER_AVERAGE

    def set_bayer_pattern(self, colorid):
        check(capi.lib.ssk_acc_set_bayer_pattern(self._h, colorid))

    def set_remap(self, rmap=None, transform=None):
        t = None if transform is None else C.byref(transform if isinstance(transform, ssk_transform) else transform.t)
        check(capi.lib.ssk_acc_set_remap(self._h, t, ref(mat(rmap))))


def compute_local_variance_map(image, dscale=1, kradius=1, uscale=0, bpp=0):
    """compute_local_variance_map (c_local_variance_sharpness_measure.cc:193-247) -> (Q, map)."""
    out = np.empty(image.shape[:2], dtype=f32)
    q = C.c_double()
    check(capi.lib.ssk_local_variance_map(C.byref(mat(np.ascontiguousarray(image))), bpp, dscale, kradius, uscale,
                                          C.byref(mat(out)), C.byref(q)))
    return q.value, out


def lpg(image, k=2.0, p=2.0, dscale=2, uscale=6):
    """lpg (core/proc/lpg.cc:223-290) -> CV_32FC1 weight map of the image size."""
    image = np.ascontiguousarray(image)
    out = np.zeros(image.shape[:2], dtype=f32)
    mo = mat(out)
    check(capi.lib.ssk_lpg(C.byref(mat(image)), float(k), float(p), int(dscale), int(uscale), C.byref(mo)))
    return out


def debayer_nn2(raw, colorid):
    """debayer_nn2 (core/io/debayer.cc:827-1195): raw Bayer HxW (uint8 / uint16 / float32) -> HxWx3 BGR of the same dtype."""
    raw = np.ascontiguousarray(raw)
    dst = np.empty(raw.shape + (3,), dtype=raw.dtype)
    check(capi.lib.ssk_debayer_nn2(C.byref(mat(raw)), C.byref(mat(dst)), int(colorid)))
    return dst


def unsharp_mask(src, sigma, alpha, outmin=-1.0, outmax=-1.0):
    """unsharp_mask (core/proc/unsharp_mask.cc:72-118) on CV_32F images; c_image_stacking_pipeline.cc:1302-1306."""
    src = np.ascontiguousarray(src, dtype=f32)
    dst = np.empty_like(src)
    check(capi.lib.ssk_unsharp_mask(C.byref(mat(src)), C.byref(mat(dst)), float(sigma), float(alpha), float(outmin), float(outmax)))
    return dst


def average_pyramid_inpaint(src, mask, max_levels=100, want_mask=True):
    """average_pyramid_inpaint (core/proc/inpaint/average_pyramid_inpaint.cc:97-127) -> (dst, dstmask)."""
    src = np.ascontiguousarray(src, dtype=f32)
    dst = np.empty_like(src)
    if mask is None:
        check(capi.lib.ssk_average_pyramid_inpaint(C.byref(mat(src)), None, C.byref(mat(dst)), None, int(max_levels)))
        return dst, None
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    dmask = np.empty_like(mask) if want_mask else None
    check(capi.lib.ssk_average_pyramid_inpaint(C.byref(mat(src)), C.byref(mat(mask)), C.byref(mat(dst)), ref(mat(dmask)),
                                               int(max_levels)))
    return dst, dmask


def gaussian_blur(src, sigma_x, sigma_y=0.0):
    """cv::GaussianBlur(src, dst, Size(), sigma_x, sigma_y, BORDER_REPLICATE) on CV_32FC1 (c_jdr_pipeline.cc:1228)."""
    src = np.ascontiguousarray(src, dtype=f32)
    out = np.zeros_like(src)
    mo = mat(out)
    check(capi.lib.ssk_gaussian_blur(C.byref(mat(src)), float(sigma_x), float(sigma_y), C.byref(mo)))
    return out


def compute_ellipsoid_zrotation_remap(size, center, axes, R1, R2, ebox_angle_deg, crop_box, wscale=1.0):
    """compute_ellipsoid_zrotation_remap (core/proc/feature2d/ellipsoid.cc:206-277); size = (w, h), crop_box = (x, y, w, h)
    -> (rmap HxWx2 float32, wmap HxW float32, rmask HxW uint8)."""
    w, h = size
    rmap = np.zeros((h, w, 2), f32)
    wmap = np.zeros((h, w), f32)
    rmask = np.zeros((h, w), np.uint8)
    mr, mw, mm = mat(rmap), mat(wmap), mat(rmask)
    d = lambda v, n: (C.c_double * n)(*[float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)])
    check(capi.lib.ssk_ellipsoid_zrotation_remap(h, w, d(center, 2), d(axes, 3), d(R1, 9), d(R2, 9), float(ebox_angle_deg),
                                                 (C.c_int * 4)(*[int(v) for v in crop_box]), float(wscale),
                                                 C.byref(mr), C.byref(mw), C.byref(mm)))
    return rmap, wmap, rmask


def jdr_derotate_and_add(acc, frame, mask, center, axes, R_current, R_target, ebox_angle_deg, crop_box, wscale, is_master,
                         enable_weighted_average=True, lpg_k=2.0, lpg_p=2.0, lpg_dscale=2, lpg_uscale=6):
    """One frame of c_jdr_pipeline::derotate_and_average_frames (c_jdr_pipeline.cc:1184-1236) into the accumulator `acc`."""
    d = lambda v, n: (C.c_double * n)(*[float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)])
    frame = np.ascontiguousarray(frame, dtype=f32)
    mm = None if mask is None else mat(np.ascontiguousarray(mask))
    check(capi.lib.ssk_jdr_derotate_and_add(acc._h, C.byref(mat(frame)), ref(mm), d(center, 2), d(axes, 3), d(R_current, 9),
                                            d(R_target, 9), float(ebox_angle_deg), (C.c_int * 4)(*[int(v) for v in crop_box]),
                                            float(wscale), int(is_master), int(enable_weighted_average), float(lpg_k),
                                            float(lpg_p), int(lpg_dscale), int(lpg_uscale)))
    return True


def set_stream_ordered(enable):
    """ssk_set_stream_ordered: device-resident lpg / gaussian_blur / acc.add / jdr_derotate_and_add calls return once enqueued
    (ordered on the device by the library).  Returns the previous mode."""
    return bool(capi.lib.ssk_set_stream_ordered(1 if enable else 0))


def device_synchronize():
    """ssk_device_synchronize: waits for everything the library has in flight."""
    check(capi.lib.ssk_device_synchronize())


def build_ellipsoid_rotation(pose):
    """build_ellipsoid_rotation(pose = (longitude_rotation, tilt_to_earth, position_angle)) (ellipsoid.h:47-71) -> 3x3 float64."""
    R = (C.c_double * 9)()
    check(capi.lib.ssk_build_ellipsoid_rotation((C.c_double * 3)(*[float(v) for v in pose]), R))
    return np.array(R, np.float64).reshape(3, 3)


def ellipsoid_bbox(size, center, axes, R):
    """ellipsoid_bbox + ellipse_crop_box (ellipsoid.cc:16-84, 279-328); size = (w, h)
    -> (((cx, cy), (width, height), angle_deg) as float32, [x, y, w, h])."""
    d = lambda v, n: (C.c_double * n)(*[float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)])
    e, cb = (C.c_float * 5)(), (C.c_int * 4)()
    check(capi.lib.ssk_ellipsoid_bbox(int(size[1]), int(size[0]), d(center, 2), d(axes, 3), d(R, 9), e, cb))
    e = np.array(e, f32)
    return ((e[0], e[1]), (e[2], e[3]), e[4]), [int(v) for v in cb]


class c_jovian_derotation_remap:
    """c_jovian_derotation_remap (core/proc/feature2d/c_jovian_derotation_remap.{h,cc}): pose bookkeeping on the host,
    compute_ellipsoid_zrotation_remap on the device."""
    default_rotation_period_sec = 9. * 3600 + 55. * 60 + 40.632       # c_jovian_derotation_remap.cc:39

    def __init__(self, rotation_period_sec=None):
        self.rotation_period_sec = self.default_rotation_period_sec if rotation_period_sec is None else rotation_period_sec
        self.rmap = self.wmap = self.rmask = None

    def set_reference_pose(self, image_size, center, axes, pose):
        self.image_size, self.center, self.axes = tuple(image_size), tuple(center), tuple(axes)
        self.target_pose = self.current_pose = tuple(float(v) for v in pose)
        self.Rtarget = self.Rcurrent = build_ellipsoid_rotation(self.target_pose)
        self.ebox, self.crop_box = ellipsoid_bbox(self.image_size, self.center, self.axes, self.Rtarget)

    def compute_derotation_for_angle(self, longitude_rotation_radians, wscale=1.0):
        self.current_pose = (self.target_pose[0] + longitude_rotation_radians, self.target_pose[1], self.target_pose[2])
        self.Rcurrent = build_ellipsoid_rotation(self.current_pose)
        self.wscale = wscale
        self.rmap, self.wmap, self.rmask = compute_ellipsoid_zrotation_remap(self.image_size, self.center, self.axes, self.Rcurrent,
                                                                            self.Rtarget, self.ebox[2], self.crop_box, wscale)

    def rotation_angle_for_time(self, deltat_sec):
        period = self.rotation_period_sec if self.rotation_period_sec > 0 else self.default_rotation_period_sec
        return 2 * np.pi * deltat_sec / period

    def compute_derotation_for_time(self, deltat_sec, wscale=1.0):
        self.compute_derotation_for_angle(self.rotation_angle_for_time(deltat_sec), wscale)

    def derotate_and_add(self, acc, frame, mask, deltat_sec, wscale, is_master, enable_weighted_average=True, **lpg):
        """One frame of derotate_and_average_frames (c_jdr_pipeline.cc:1184-1236 / c_sdr_pipeline.cc:1192-1246): the caller's
        compute_derotation_for_time(-dt, w) and everything after it, fused on the device."""
        angle = self.rotation_angle_for_time(deltat_sec)
        self.current_pose = (self.target_pose[0] + angle, self.target_pose[1], self.target_pose[2])
        self.Rcurrent = build_ellipsoid_rotation(self.current_pose)
        return jdr_derotate_and_add(acc, frame, mask, self.center, self.axes, self.Rcurrent, self.Rtarget, self.ebox[2], self.crop_box,
                                    wscale, is_master, enable_weighted_average, **lpg)


class c_saturn_derotation_remap(c_jovian_derotation_remap):
    """c_saturn_derotation_remap (core/proc/feature2d/c_saturn_derotation_remap.{h,cc}): the Jovian class with Saturn's
    System III period, as in the reference."""
    default_rotation_period_sec = 10 * 3600. + 33 * 60. + 38          # c_saturn_derotation_remap.cc:34


def linear_interpolation_inpaint(src, mask):
    """linear_interpolation_inpaint (core/proc/inpaint/linear_interpolation_inpaint.cc:327-368) on CV_32F images -> filled copy."""
    src = np.ascontiguousarray(src, dtype=f32)
    dst = np.empty_like(src)
    mm = None if mask is None else mat(np.ascontiguousarray(mask, dtype=np.uint8))
    check(capi.lib.ssk_linear_interpolation_inpaint(C.byref(mat(src)), ref(mm), C.byref(mat(dst))))
    return dst


def median_filter_bad_pixels(image, variation_threshold):
    """median_filter_bad_pixels (core/proc/bad_pixels.cc:58-70) of a mono / colour frame -> filtered copy."""
    out = np.ascontiguousarray(image).copy()
    m = mat(out)
    check(capi.lib.ssk_median_filter_bad_pixels(C.byref(m), float(variation_threshold)))
    return out


def bayer_denoise(raw, variation_threshold):
    """bayer_denoise (core/io/debayer.cc:1471-1611, returnBayerPlanes = false) of a raw single-channel Bayer frame -> filtered copy."""
    out = np.ascontiguousarray(raw).copy()
    m = mat(out)
    check(capi.lib.ssk_bayer_denoise(C.byref(m), float(variation_threshold)))
    return out


def average_bayer_planes(raw):
    """average_bayer_planes (core/io/debayer.cc:277-376) of a raw single-channel Bayer frame -> half-size image of the same dtype."""
    raw = np.ascontiguousarray(raw)
    dst = np.empty((raw.shape[0] // 2, raw.shape[1] // 2), dtype=raw.dtype)
    check(capi.lib.ssk_average_bayer_planes(C.byref(mat(raw)), C.byref(mat(dst))))
    return dst


def input_calibrate(frame, bpp=0, dark=None, flat=None):
    """read_input_frame's dark / flat correction (c_image_stacking_pipeline_base.cc:143-184) -> CV_32F frame."""
    frame = np.ascontiguousarray(frame)
    dst = np.empty(frame.shape, dtype=f32)
    md = None if dark is None else mat(np.ascontiguousarray(dark, dtype=f32))
    mf = None if flat is None else mat(np.ascontiguousarray(flat, dtype=f32))
    check(capi.lib.ssk_input_calibrate(C.byref(mat(frame)), int(bpp), ref(md), ref(mf), C.byref(mat(dst))))
    return dst


def color_transform(image, matrix):
    """cv::transform(image, image, color_matrix) (c_image_stacking_pipeline_base.cc:263-266) on CV_32FC3 images."""
    image = np.ascontiguousarray(image, dtype=f32)
    m = np.ascontiguousarray(matrix, dtype=f32)
    assert m.shape in ((3, 3), (3, 4))
    dst = np.empty_like(image)
    check(capi.lib.ssk_color_transform(C.byref(mat(image)), m.ctypes.data_as(C.POINTER(C.c_float)), m.shape[1], C.byref(mat(dst))))
    return dst


class c_ser_reader:
    """c_ser_reader (core/io/c_ser_file.h:136-190): open / seek / read of a SER sequence."""

    def __init__(self, path):
        self._h = C.c_void_p()
        check(capi.lib.ssk_ser_open(os.fsencode(path), C.byref(self._h)))
        v = [C.c_int() for _ in range(7)]
        check(capi.lib.ssk_ser_info(self._h, *[C.byref(x) for x in v]))
        self.cols, self.rows, self.type, self.bits_per_plane, self.color_id, self.num_frames, ts = [x.value for x in v]
        self.has_timestamps = bool(ts)
        depth, cn = self.type & 7, (self.type >> 3) + 1
        self._dtype = {capi.SSK_8U: np.uint8, capi.SSK_16U: np.uint16, capi.SSK_32F: np.float32}[depth]
        self._shape = (self.rows, self.cols) if cn == 1 else (self.rows, self.cols, cn)

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_ser_close(self._h)
            self._h = None

    def bpp(self):
        """c_input_source::bpp(): bits per sample of integer frames (the 1 / (1 << bpp) scale of read_input_frame)."""
        return self.bits_per_plane if self.bits_per_plane > 0 else 0

    def read(self, index, out=None):
        """-> (frame, timestamp).  `out`: an array to read into (e.g. a pinned buffer)."""
        img = np.empty(self._shape, dtype=self._dtype) if out is None else out
        ts = C.c_uint64(0)
        check(capi.lib.ssk_ser_read(self._h, int(index), C.byref(mat(img)), C.byref(ts)))
        return img, ts.value


def select_master_frame(frames, colorid=capi.COLORID_MONO, dscale=1, kradius=1, uscale=0):
    """The master_frame_best_of_100_in_middle branch of select_master_frame (c_image_stacking_pipeline_base.cc:311-399) over the
    scanned frames: the sharpness metric of compute_local_variance_map on each frame (on average_bayer_planes of a raw Bayer
    frame), first maximum wins.  -> (best_index, metrics)."""
    best, best_metric, metrics = 0, 0.0, []
    for i, f in enumerate(frames):
        tmp = average_bayer_planes(f) if colorid in (capi.COLORID_BAYER_RGGB, capi.COLORID_BAYER_GRBG, capi.COLORID_BAYER_GBRG,
                                                      capi.COLORID_BAYER_BGGR) else f
        q = C.c_double()
        check(capi.lib.ssk_local_variance_map(C.byref(mat(np.ascontiguousarray(tmp))), -1, dscale, kradius, uscale, None, C.byref(q)))
        metrics.append(q.value)
        if q.value > best_metric:
            best_metric, best = q.value, i
    return best, metrics


def master_frame_range(num_frames, master_frame_pos, max_frames_to_stack, start_frame_index=0):
    """[startpos, endpos) of the frames create_reference_frame stacks into the master frame (c_image_stacking_pipeline.cc:1211-1229)."""
    start_frame_index = max(0, start_frame_index)
    if start_frame_index + max_frames_to_stack >= num_frames:
        return start_frame_index, num_frames
    startpos = max(start_frame_index, master_frame_pos - max_frames_to_stack // 2)
    endpos = startpos + max_frames_to_stack
    if endpos >= num_frames:
        startpos = max(start_frame_index, num_frames - max_frames_to_stack)
        endpos = num_frames
    return startpos, endpos


def create_reference_frame(frames, master_frame_pos, options, max_frames_to_stack=3000, unsharp_sigma=1.0, unsharp_alpha=0.8, bpp=0):
    """c_image_stacking_pipeline::create_reference_frame (c_image_stacking_pipeline.cc:1112-1312) for an in-memory sequence:
    the master frame alone (max_frames_to_stack < 2), or the stack of the frames around it registered against it with
    BORDER_REFLECT101 (generating_master_frame), compute(), linear_interpolation_inpaint; then unsharp_mask(sigma, alpha).
    `options`: ssk_stack_options of the master pass.  -> (reference_frame CV_32F, reference_mask or None)."""
    frames = list(frames)
    ref0 = frames[master_frame_pos]
    scale = 1.0 if ref0.dtype == np.float32 else 1.0 / (1 << bpp)
    reference, mask = (ref0.astype(np.float64) * scale).astype(f32) if ref0.dtype != np.float32 else ref0.copy(), None
    if max_frames_to_stack >= 2 and len(frames) >= 2:
        o = capi.ssk_stack_options.from_buffer_copy(options)
        o.generating_master_frame = 1
        p = c_image_stacking_pipeline(o)
        p.set_reference(ref0, bpp=bpp)
        lo, hi = master_frame_range(len(frames), master_frame_pos, max_frames_to_stack)
        mb = max(1, int(o.max_batch))
        for i in range(lo, hi, mb):
            p.add_frames(frames[i:min(hi, i + mb)], want_results=False)
        if p.accumulated_frames() < 1:
            raise RuntimeError("No frames accumulated for reference frame")
        reference, mask = p.compute()
        reference = linear_interpolation_inpaint(reference, mask)
    if unsharp_sigma > 0 and unsharp_alpha > 0:
        reference = unsharp_mask(reference, unsharp_sigma, unsharp_alpha)
    return reference, mask


def stack_options(**kw):
    o = capi.ssk_stack_options()
    capi.lib.ssk_stack_options_default(C.byref(o))
    reg = kw.pop("registration", None)
    if reg is not None:
        o.registration = reg
    for k, v in kw.items():
        assert hasattr(o, k), k
        setattr(o, k, v)
    return o


class c_image_stacking_pipeline:
    """The per-frame loop of c_image_stacking_pipeline::process_input_sequence
    (c_image_stacking_pipeline.cc:1358-1862) + finalise (:731-769), batched on the device."""

    def __init__(self, options):
        self.options = options
        self._h = C.c_void_p()
        check(capi.lib.ssk_stack_create(C.byref(options), C.byref(self._h)))
        self._shape = None

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None:
            capi.lib.ssk_stack_destroy(self._h)
            self._h = None

    def set_reference(self, image, bpp=0, mask=None):
        m = image if isinstance(image, ssk_mat) else mat(np.ascontiguousarray(image))
        mm = mask if (mask is None or isinstance(mask, ssk_mat)) else mat(np.ascontiguousarray(mask))
        check(capi.lib.ssk_stack_set_reference(self._h, C.byref(m), ref(mm), bpp))
        bayer = self.options.accumulation_method == capi.STACK_BAYER_AVERAGE
        rows, cols = m.rows, m.cols
        o = self.options
        if o.upscale_option and o.enable_registration and not o.generating_master_frame:   # frame_upscale_after_align
            uc, ur = C.c_int(), C.c_int()
            check(capi.lib.ssk_upscale_size(o.upscale_option, cols, rows, C.byref(uc), C.byref(ur)))
            rows, cols = ur.value, uc.value
        self._shape = (rows, cols, 3 if bayer else (m.type >> 3) + 1)   # c_bayer_average computes a BGR image
        self._bpp = bpp

    def _mats(self, frames):
        arr = (ssk_mat * len(frames))()
        keep = []
        for i, f in enumerate(frames):
            m = f if isinstance(f, ssk_mat) else mat(np.ascontiguousarray(f))
            keep.append(m)
            arr[i] = m
        return arr, keep

    def add_frames(self, frames, want_results=True):
        arr, keep = self._mats(frames)
        n = len(frames)
        ts = (ssk_transform * n)() if want_results else None
        st = (ssk_ecc_status * n)() if want_results else None
        check(capi.lib.ssk_stack_add_frames(self._h, arr, n, self._bpp, ts, st))
        if not want_results:
            return None
        return [dict(ok=bool(st[i].ok), params=ts[i].parameters(), rho=st[i].rho, eps=st[i].eps,
                     iterations=st[i].num_iterations) for i in range(n)]

    def run_stacking_pass(self, frames, reference, bpp=0, unsharp_sigma=1.0, unsharp_alpha=0.8, inpaint_max_levels=100):
        """One stacking pass with the steps either side of the per-frame loop, as run_pipeline does them:
        unsharp_mask of the master / reference frame (c_image_stacking_pipeline.cc:1302-1306), set_reference, the
        batched per-frame loop, compute() + average_pyramid_inpaint (c_image_stacking_pipeline.cc:742-767).
        `reference` is CV_32F (a master frame is always float); returns (avg, mask, per-frame results)."""
        ref_img = np.ascontiguousarray(reference, dtype=f32)
        if unsharp_sigma > 0 and unsharp_alpha > 0:
            ref_img = unsharp_mask(ref_img, unsharp_sigma, unsharp_alpha)
        self.set_reference(ref_img, bpp=bpp)      # a float reference is never rescaled; bpp applies to integer frames
        res = []
        mb = max(1, int(self.options.max_batch))
        frames = list(frames)
        for i in range(0, len(frames), mb):
            res += self.add_frames(frames[i:i + mb])
        avg, mask = self.compute(inpaint_max_levels=inpaint_max_levels)
        return avg, mask, res

    def submit(self, frames):
        """Streaming form: enqueue up to max_batch frames, return a ticket at once (ssk_stack_submit)."""
        arr, keep = self._mats(frames)
        t = C.c_int64(-1)
        check(capi.lib.ssk_stack_submit(self._h, arr, len(frames), self._bpp, C.byref(t)))
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[t.value] = (keep, frames)      # host frames must outlive the upload
        return t.value

    def wait(self, ticket):
        """Results of the chunk `ticket` (ssk_stack_wait)."""
        n = len(self._inflight[ticket][1])
        ts, st, m = (ssk_transform * n)(), (ssk_ecc_status * n)(), C.c_int(0)
        check(capi.lib.ssk_stack_wait(self._h, ticket, ts, st, n, C.byref(m)))
        del self._inflight[ticket]
        return [dict(ok=bool(st[i].ok), params=ts[i].parameters(), rho=st[i].rho, eps=st[i].eps,
                     iterations=st[i].num_iterations) for i in range(m.value)]

    def add_frames_async(self, frames):
        arr, keep = self._mats(frames)
        check(capi.lib.ssk_stack_add_frames_async(self._h, arr, len(frames), self._bpp))
        return keep

    def sync(self):
        check(capi.lib.ssk_stack_sync(self._h))

    def reset(self):
        """Empty accumulator and frame count for a new run over the same reference (ssk_stack_reset)."""
        check(capi.lib.ssk_stack_reset(self._h))

    def flush(self):
        """Stream-side join of the ring kernel left running by the last device-frame call (no host sync)."""
        check(capi.lib.ssk_stack_flush(self._h))

    def accumulated_frames(self):
        return capi.lib.ssk_stack_accumulated_frames(self._h)

    def compute(self, inpaint_max_levels=None):
        """(avg, mask) of the accumulator; with inpaint_max_levels the holes are filled on the device by
        average_pyramid_inpaint, as c_image_stacking_pipeline.cc:742-767 does at the end of a run."""
        h, w, c = self._shape
        avg = np.empty((h, w) if c == 1 else (h, w, c), dtype=f32)
        mask = np.empty((h, w), dtype=np.uint8)
        if inpaint_max_levels is None:
            check(capi.lib.ssk_stack_compute(self._h, C.byref(mat(avg)), C.byref(mat(mask))))
        else:   # c_image_stacking_pipeline.cc:763-767 passes 100
            check(capi.lib.ssk_stack_compute_inpainted(self._h, C.byref(mat(avg)), C.byref(mat(mask)), int(inpaint_max_levels)))
        return avg, mask

    def accumulator(self):
        if self.options.accumulation_method == capi.STACK_BAYER_AVERAGE:
            return c_bayer_average(capi.lib.ssk_stack_accumulator(self._h))
        return c_weigthed_average(capi.lib.ssk_stack_accumulator(self._h))

    def stream(self):
        return capi.lib.ssk_stack_stream(self._h)

    def stage_times(self):
        ms = (C.c_float * 4)()
        check(capi.lib.ssk_stack_stage_times(self._h, ms))
        return list(ms)
