"""
Oracle restatement of the per-frame stacking loop.

Follows /root/reference/core/pipeline/c_image_stacking_pipeline/c_image_stacking_pipeline.cc:
  create_frame_accumulation   :450-466
  process_input_sequence      :1358-1862  (ordering: weights -> register -> remap(frame, mask) ->
                                           remap(weights) -> weights*mask -> add)
  multiply_weights            :108-138, 1704-1714
  weights_required            :2004-2011, compute_weights :2013-2020
  finalise (compute)          :731-769 (average_pyramid_inpaint: oracle/inpaint.py)
and the input scaling of c_image_stacking_pipeline_base.cc:271-276 (convertTo(CV_32F, 1/(1<<bpp))).

Test infrastructure only (see oracle/__init__.py).
"""
from dataclasses import dataclass, field
import numpy as np
import cv2

from .registration import FrameRegistration, ImageRegistrationOptions
from .accumulation import WeightedAverage, BayerAverage
from .weights import compute_local_variance_map

f32 = np.float32

ACC_AVERAGE = "average"
ACC_WEIGHTED_AVERAGE = "weighted_average"
ACC_BAYER_AVERAGE = "bayer_average"


@dataclass
class StackingOptions:
    registration: ImageRegistrationOptions = field(default_factory=ImageRegistrationOptions)
    accumulation_method: str = ACC_AVERAGE
    # c_frame_accumulation_options::sharpness_measure (c_image_stacking_pipeline.h:94-99)
    sm_dscale: int = 1
    sm_kradius: int = 1
    sm_uscale: int = 0
    enable_registration: bool = True
    # c_frame_upscale_options (c_image_stacking_pipeline.h:33-86); only frame_upscale_after_align is restated
    upscale_option: int = 0      # 0 none, 1 x2.0 (pyrUp), 2 x1.5, 3 x3.0
    generating_master_frame: bool = False


def to_float_frame(raw, bpp):
    """c_image_stacking_pipeline_base.cc:271-276."""
    if raw.dtype == np.float32:
        return raw
    return (raw.astype(np.float64) * (1.0 / (1 << bpp))).astype(f32)


UPSCALE_NONE, UPSCALE_PYRUP, UPSCALE_X15, UPSCALE_X30 = 0, 1, 2, 3


def _resize_no_ipp(src, dsize, interpolation):
    """cv::resize as a distribution build of OpenCV computes it (the reference links the system library, CMakeLists.txt:47-59):
    the IPP code path of the cv2 wheel uses another arithmetic for INTER_LINEAR."""
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        return cv2.resize(src, dsize, interpolation=interpolation)
    finally:
        cv2.ipp.setUseIPP(was)


def upscale_remap(option, srcmap):
    """c_image_stacking_pipeline::upscale_remap (c_image_stacking_pipeline.cc:1869-1905)."""
    h, w = srcmap.shape[:2]
    if option == UPSCALE_X15:
        return _resize_no_ipp(srcmap, (w * 3 // 2, h * 3 // 2), cv2.INTER_LINEAR)
    if option == UPSCALE_PYRUP:
        return cv2.pyrUp(srcmap)
    if option == UPSCALE_X30:
        return _resize_no_ipp(srcmap, (w * 3, h * 3), cv2.INTER_LINEAR_EXACT)
    return srcmap.copy()


def upscale_optflow(option, srcmap):
    """c_image_stacking_pipeline::upscale_optflow (c_image_stacking_pipeline.cc:1907-1946)."""
    f = {UPSCALE_X15: 1.5, UPSCALE_PYRUP: 2.0, UPSCALE_X30: 3.0}.get(option)
    dst = upscale_remap(option, srcmap)
    return dst if f is None else cv2.multiply(dst, f)


def upscale_image(option, src, srcmask=None):
    """c_image_stacking_pipeline::upscale_image (c_image_stacking_pipeline.cc:1949-2000) -> (dst, dstmask)."""
    if option == UPSCALE_NONE:
        return (None if src is None else src.copy()), (None if srcmask is None else srcmask.copy())
    def up(img):
        h, w = img.shape[:2]
        if option == UPSCALE_X15:
            return _resize_no_ipp(img, (w * 3 // 2, h * 3 // 2), cv2.INTER_LINEAR)
        if option == UPSCALE_PYRUP:
            return cv2.pyrUp(img)
        return _resize_no_ipp(img, (w * 3, h * 3), cv2.INTER_LINEAR_EXACT)
    dst = None if src is None else up(src)
    dmask = None
    if srcmask is not None:
        dmask = cv2.compare(up(srcmask), 255, cv2.CMP_GE)
    return dst, dmask


def weights_required(o: StackingOptions):
    return o.accumulation_method == ACC_WEIGHTED_AVERAGE and o.sm_kradius > 0


def process_frame(reg: FrameRegistration, acc, o: StackingOptions, frame, mask=None, raw_bayer=None):
    """One iteration of process_input_sequence. Returns True if the frame was accumulated."""
    weights = None
    if weights_required(o):
        _, weights = compute_local_variance_map(frame, o.sm_dscale, o.sm_kradius, o.sm_uscale)
        if weights is not None and mask is not None:
            weights[mask == 0] = 0
    ro = o.registration
    if o.enable_registration:
        if not reg.register_frame(frame, mask):
            return False            # c_image_stacking_pipeline.cc:1578-1581: frame dropped
        rmap = reg.current_remap
        if o.upscale_option != UPSCALE_NONE and not o.generating_master_frame:
            rmap = upscale_remap(o.upscale_option, rmap)      # c_image_stacking_pipeline.cc:1633-1642
        frame, mask = reg.custom_remap(rmap, frame, mask, ro.interpolation, ro.border_mode, ro.border_value)
        if weights is not None:
            weights, _ = reg.custom_remap(rmap, weights, None, ro.interpolation, cv2.BORDER_CONSTANT, None,
                                          want_mask=False)
    if weights is not None:
        if mask is not None:
            weights = cv2.multiply(mask, weights, scale=(1.0 / 255 if mask.dtype == np.uint8 else 1.0),
                                   dtype=cv2.CV_32F)
        mask = weights
    if isinstance(acc, BayerAverage):
        acc.set_remap(reg.current_remap if o.enable_registration else None)
        return acc.add(raw_bayer, mask)
    return acc.add(frame, mask)


def run_stacking(frames, o: StackingOptions, reference=None, collect=None):
    """frames: iterable of CV_32F frames. reference defaults to frames[0].
    Returns (avg, mask, accumulator, registration); per-frame records appended to `collect` if given."""
    frames = list(frames)
    reg = FrameRegistration(o.registration)
    ref = frames[0] if reference is None else reference
    if o.enable_registration:
        reg.setup_reference_frame(ref, None)
    acc = BayerAverage() if o.accumulation_method == ACC_BAYER_AVERAGE else WeightedAverage()
    for f in frames:
        ok = process_frame(reg, acc, o, f)
        if collect is not None:
            collect.append(dict(ok=ok, params=None if reg.image_transform is None else reg.image_transform.clone_parameters(),
                                rho=reg.status.rho, eps=reg.status.eps, iterations=reg.status.num_iterations))
    avg, mask = acc.compute()
    return avg, mask, acc, reg


def run_bayer_stacking(raw_frames, bpp, o: StackingOptions, colorid, reference=None, collect=None):
    """The bayer_average form of the loop on raw Bayer frames (config #3): read_input_frame keeps the raw samples as
    _raw_bayer_image (CV_32F, 1/(1<<bpp)), demosaics the frame with debayer_nn2 at its own depth and converts the result
    to CV_32F (c_image_stacking_pipeline_base.cc:221-236, 271-276); the registration sees the demosaiced frame, the
    accumulator gathers the raw samples through current_remap under the remapped mask
    (c_image_stacking_pipeline.cc:1644-1651, 1730-1752).  reference: raw Bayer frame (default raw_frames[0]).
    Returns (avg HxWx3, mask, accumulator, registration)."""
    from .debayer import debayer_nn2
    assert o.accumulation_method == ACC_BAYER_AVERAGE
    raw_frames = list(raw_frames)
    reg = FrameRegistration(o.registration)
    ref = raw_frames[0] if reference is None else reference
    if o.enable_registration:
        reg.setup_reference_frame(to_float_frame(debayer_nn2(ref, colorid), bpp), None)
    acc = BayerAverage()
    acc.set_bayer_pattern(colorid)
    for f in raw_frames:
        ok = process_frame(reg, acc, o, to_float_frame(debayer_nn2(f, colorid), bpp), None, raw_bayer=to_float_frame(f, bpp))
        if collect is not None:
            collect.append(dict(ok=ok, params=None if reg.image_transform is None else reg.image_transform.clone_parameters(),
                                rho=reg.status.rho, eps=reg.status.eps, iterations=reg.status.num_iterations))
    avg, mask = acc.compute()
    return avg, mask, acc, reg


def run_stacking_pass(frames, o: StackingOptions, reference, unsharp_sigma=1.0, unsharp_alpha=0.8, inpaint_max_levels=100,
                      collect=None):
    """The stacking pass with the steps either side of the per-frame loop:
      * the master / reference frame is sharpened by unsharp_mask(sigma, alpha) when both are positive
        (c_image_stacking_pipeline.cc:1302-1306, defaults c_image_stacking_pipeline.h:169-170),
      * the accumulator read-out is finished by average_pyramid_inpaint(avg, mask, avg, mask, 100)
        (c_image_stacking_pipeline.cc:742-767).
    Returns (avg, mask, sharpened_reference)."""
    from .unsharp import unsharp_mask
    from .inpaint import average_pyramid_inpaint
    ref = np.ascontiguousarray(reference, dtype=f32)
    if unsharp_sigma > 0 and unsharp_alpha > 0:
        ref = unsharp_mask(ref, unsharp_sigma, unsharp_alpha)
    avg, mask, _, _ = run_stacking(frames, o, reference=ref, collect=collect)
    avg, mask = average_pyramid_inpaint(avg, mask, inpaint_max_levels)
    return avg, mask, ref


def master_frame_range(num_frames, master_frame_pos, max_frames_to_stack, start_frame_index=0):
    """[startpos, endpos) of create_reference_frame (c_image_stacking_pipeline.cc:1211-1229)."""
    start_frame_index = max(0, start_frame_index)
    if start_frame_index + max_frames_to_stack >= num_frames:
        return start_frame_index, num_frames
    startpos = max(start_frame_index, master_frame_pos - max_frames_to_stack // 2)
    endpos = startpos + max_frames_to_stack
    if endpos >= num_frames:
        startpos = max(start_frame_index, num_frames - max_frames_to_stack)
        endpos = num_frames
    return startpos, endpos


def create_reference_frame(frames, master_frame_pos, o: StackingOptions, max_frames_to_stack=3000, unsharp_sigma=1.0,
                           unsharp_alpha=0.8):
    """c_image_stacking_pipeline::create_reference_frame (c_image_stacking_pipeline.cc:1112-1312) for CV_32F frames: the master
    pass stacks the frames around the selected one registered against it (remap border REFLECT101 while generating the master
    frame, :1644-1651), then compute(), linear_interpolation_inpaint, unsharp_mask."""
    import copy
    from .inpaint import linear_interpolation_inpaint
    from .unsharp import unsharp_mask
    frames = list(frames)
    reference, mask = frames[master_frame_pos].copy(), None
    if max_frames_to_stack >= 2 and len(frames) >= 2:
        om = copy.deepcopy(o)
        om.registration.border_mode = cv2.BORDER_REFLECT101
        lo, hi = master_frame_range(len(frames), master_frame_pos, max_frames_to_stack)
        reference, mask, _, _ = run_stacking(frames[lo:hi], om, reference=reference)
        reference = linear_interpolation_inpaint(reference, mask)
    if unsharp_sigma > 0 and unsharp_alpha > 0:
        reference = unsharp_mask(reference, unsharp_sigma, unsharp_alpha)
    return reference, mask


def select_master_frame(frames, bayer=False, dscale=1, kradius=1, uscale=0):
    """master_frame_best_of_100_in_middle over the scanned frames (c_image_stacking_pipeline_base.cc:311-399) -> (best, metrics)."""
    from .debayer import average_bayer_planes
    best, best_metric, metrics = 0, 0.0, []
    for i, f in enumerate(frames):
        tmp = average_bayer_planes(f) if bayer else f
        q, _ = compute_local_variance_map(tmp, dscale, kradius, uscale, full_resolution=False)
        metrics.append(q)
        if q > best_metric:
            best_metric, best = q, i
    return best, metrics
