"""
Oracle restatement of c_frame_registration (ECC branch only; the sparse-feature branch is out of scope).

Follows /root/reference/core/proc/image_registration/c_frame_registration.{h,cc}:
  options structs        c_frame_registration.h:47-64, 119-136
  scaleImage             c_frame_registration.cc:230-250
  create_image_transfrom c_frame_registration.cc:383-407
  setup_reference_frame  c_frame_registration.cc:565-719
  register_frame         c_frame_registration.cc:721-964
  create_ecc_image       c_frame_registration.cc:1009-1022 (extract_channel gray: extract_channel.cc:546-575, 607-697)
  base_remap/custom_remap/remap  c_frame_registration.cc:1265-1417

Test infrastructure only (see oracle/__init__.py).
"""
from dataclasses import dataclass, field
import numpy as np
import cv2

from . import ecc as _ecc
from . import eccflow as _flow
from . import transforms as _tf

f32 = np.float32


@dataclass
class EccRegistrationOptions:
    """c_ecc_registration_options (c_frame_registration.h:47-64), same defaults."""
    scale: float = 0.5
    eps: float = 0.2
    min_rho: float = 0.8
    input_smooth_sigma: float = 1.0
    reference_smooth_sigma: float = 1.0
    update_step_scale: float = 1.5
    se_radius: int = 5
    ecc_method: int = _ecc.ECC_ALIGN_LM
    max_iterations: int = 50
    ecch_max_level: int = 0
    ecch_minimum_image_size: int = 16
    normalization_noise: float = 0.01
    normalization_scale: int = 0
    ecch_estimate_translation_first: bool = True
    replace_planetary_disk_with_mask: bool = False


@dataclass
class ImageRegistrationOptions:
    """c_image_registration_options (c_frame_registration.h:119-136); feature/eccflow members omitted."""
    motion_type: int = _tf.IMAGE_MOTION_AFFINE
    interpolation: int = cv2.INTER_LINEAR
    border_mode: int = cv2.BORDER_REFLECT101
    border_value: tuple = (0.0, 0.0, 0.0, 0.0)
    ecc: EccRegistrationOptions = field(default_factory=EccRegistrationOptions)
    enable_feature_registration: bool = False   # reference default is True; out of scope here
    enable_ecc_registration: bool = True        # reference default is False
    enable_eccflow_registration: bool = False
    eccflow: object = None                      # c_eccflow_registration_options (oracle.eccflow.registration_options())


@dataclass
class EccStatus:
    rho: float = 0.0
    eps: float = 0.0
    num_iterations: int = 0
    ok: bool = False


def scale_image(scale, src, srcm):
    """c_frame_registration.cc:230-250."""
    if abs(scale - 0.5) < 1e-2:
        dst = cv2.pyrDown(src)
        dstm = None
        if srcm is not None:
            dstm = cv2.pyrDown(srcm, dstsize=(dst.shape[1], dst.shape[0]))
            dstm = cv2.compare(dstm, 250, cv2.CMP_GE)
    else:
        dst = cv2.resize(src, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_AREA)
        dstm = None
        if srcm is not None:
            dstm = cv2.resize(srcm, (dst.shape[1], dst.shape[0]), interpolation=cv2.INTER_AREA)
            dstm = cv2.compare(dstm, 250, cv2.CMP_GE)
    return dst, dstm


def create_ecc_image(src, srcm):
    """c_frame_registration.cc:1009-1022 with ecc_registration_channel = gray, ddepth CV_32F, autoscale.

    Inside the stacking pipeline frames are already CV_32F in [0,1) (c_image_stacking_pipeline_base.cc:271-276),
    so the depth conversion is a copy; colour goes through cv::cvtColor(COLOR_BGR2GRAY)."""
    assert src.dtype == np.float32, "pipeline frames are CV_32F when they reach registration"
    if src.ndim == 3 and src.shape[2] > 1:
        dst = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
    else:
        dst = src.reshape(src.shape[:2]).copy()
    dstm = srcm
    if srcm is not None and srcm.ndim == 3:
        dstm = srcm.max(axis=2)
    return dst, dstm


class FrameRegistration:
    """c_frame_registration, ECC branch (c_frame_registration.h:210-354)."""

    def __init__(self, options: ImageRegistrationOptions):
        self.options = options
        self.ecch = _ecc.EccH()
        self.image_transform = None
        self._default_parameters = None
        self.current_remap = None
        self.reference_frame_size = None
        self.status = EccStatus()

    def _create_image_transform(self):
        # c_frame_registration.cc:383-407
        if self.image_transform is None:
            self.ecch.set_image_transform(None)
            self.image_transform = _tf.create_image_transform(self.options.motion_type)
            self._default_parameters = self.image_transform.clone_parameters()
            if self.options.enable_ecc_registration and self.options.ecc.scale > 0:
                self.ecch.set_image_transform(self.image_transform)
        return True

    def setup_reference_frame(self, reference_image, reference_mask=None):
        # c_frame_registration.cc:565-719
        o = self.options
        self.reference_frame_size = (reference_image.shape[1], reference_image.shape[0])
        ref_ecc_image, ref_ecc_mask = create_ecc_image(reference_image, reference_mask)
        e = self.ecch
        e.method = o.ecc.ecc_method
        e.epsx = o.ecc.eps
        e.input_smooth_sigma = o.ecc.input_smooth_sigma
        e.reference_smooth_sigma = o.ecc.reference_smooth_sigma
        e.update_step_scale = o.ecc.update_step_scale
        e.max_iterations = o.ecc.max_iterations
        e.maxlevel = o.ecc.ecch_max_level
        e.minimum_image_size = o.ecc.ecch_minimum_image_size
        e.pyramid = []  # set_method/set_maxlevel/set_minimum_image_size clear the pyramid (ecc2.cc:739-767)
        if o.ecc.scale > 0 and o.ecc.scale != 1:
            ecc_image, ecc_mask = scale_image(o.ecc.scale, ref_ecc_image, ref_ecc_mask)
        else:
            ecc_image, ecc_mask = ref_ecc_image, ref_ecc_mask
        if o.ecc.normalization_scale > 0 and o.ecc.normalization_noise > 0:
            ecc_image = _ecc.ecc_normalize(ecc_image, ecc_mask, o.ecc.normalization_scale)
        if not e.set_reference_image(ecc_image, ecc_mask):
            return False
        if o.enable_eccflow_registration:
            # c_frame_registration.cc:637-660: the flow works on the full-size ECC image and mask
            self.eccflow = _flow.EccFlow(o.eccflow if o.eccflow is not None else _flow.registration_options())
            return self.eccflow.set_reference_image(ref_ecc_image, ref_ecc_mask)
        return True

    def register_frame(self, current_image, current_mask=None):
        # c_frame_registration.cc:721-964 (dst/dstmask not requested, as in the stacking pipeline)
        o = self.options
        self.status = EccStatus()
        self._create_image_transform()
        t = self.image_transform
        t.set_parameters(self._default_parameters)
        ecc_image, ecc_mask = create_ecc_image(current_image, current_mask)
        if o.ecc.scale > 0 and o.ecc.scale != 1:
            cur_img, cur_mask = scale_image(o.ecc.scale, ecc_image, ecc_mask)
        else:
            cur_img, cur_mask = ecc_image, ecc_mask
        if o.ecc.normalization_scale > 0 and o.ecc.normalization_noise > 0:
            cur_img = _ecc.ecc_normalize(cur_img, cur_mask, o.ecc.normalization_scale)

        estimate_translation_first = (o.motion_type != _tf.IMAGE_MOTION_TRANSLATION and
                                      o.ecc.ecch_estimate_translation_first and o.ecc.ecch_max_level != 0)
        e = self.ecch
        if estimate_translation_first:
            tt = _tf.TranslationTransform(*t.translation())
            e.set_image_transform(tt)
            e.align(cur_img, cur_mask)
            rho = _ecc.compute_correlation(e.current_image(), e.current_mask(), e.reference_image(),
                                           e.reference_mask(), e.create_remap())
            self.status.rho = rho
            if rho < 0.75 * o.ecc.min_rho:
                e.set_image_transform(t)
                return False
            t.set_translation(tt.translation())
            e.set_image_transform(t)

        e.align(cur_img, cur_mask)
        rho = _ecc.compute_correlation(e.current_image(), e.current_mask(), e.reference_image(),
                                       e.reference_mask(), e.create_remap())
        self.status.rho = rho
        self.status.eps = e.eps()
        self.status.num_iterations = e.num_iterations
        if rho < o.ecc.min_rho:
            return False
        if o.ecc.scale > 0 and o.ecc.scale != 1:
            t.scale_transfrom(1.0 / o.ecc.scale)
        self.current_remap = t.create_remap(self.reference_frame_size)
        if o.enable_eccflow_registration:
            # c_frame_registration.cc:900-917
            self.current_remap = self.eccflow.compute(ecc_image, self.current_remap, ecc_mask)
        self.status.ok = True
        return True

    def base_remap(self, rmap, src, src_mask, interpolation=None, border_mode=None, border_value=None,
                   want_dst=True, want_mask=True):
        """c_frame_registration.cc:1265-1386 -> (dst, dst_mask)."""
        o = self.options
        if interpolation is None or interpolation < 0:
            interpolation = o.interpolation
        if border_mode is not None and border_mode >= 0:
            bv = (0.0, 0.0, 0.0, 0.0) if border_value is None else border_value
        else:
            border_mode = o.border_mode
            bv = o.border_value
        dst = dst_mask = None
        if want_dst:
            dst = cv2.remap(src, rmap, None, interpolation, borderMode=border_mode, borderValue=bv)
        if want_mask:
            if src_mask is not None:
                m = src_mask
            else:
                size = src.shape[:2] if src is not None else rmap.shape[:2]
                m = np.full(size, 255, dtype=np.uint8)
            dst_mask = cv2.remap(m, rmap, None, interpolation, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
            if dst_mask.dtype == np.uint8:
                dst_mask = cv2.compare(dst_mask, 255, cv2.CMP_GE)
                dst_mask = cv2.erode(dst_mask, np.full((5, 5), 255, np.uint8), borderType=cv2.BORDER_CONSTANT,
                                     borderValue=(255, 255, 255, 255))
        return dst, dst_mask

    custom_remap = base_remap

    def remap(self, src, src_mask, **kw):
        return self.base_remap(self.current_remap, src, src_mask, **kw)
