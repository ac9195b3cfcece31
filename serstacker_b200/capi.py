"""
ctypes binding of include/ssk.h (serstacker_b200/libssk.so).

This is the only way Python reaches the CUDA code: plain pointers and sizes through the C ABI, no torch types.
Importing this module fails loudly when the library has not been built; every compute entry point fails with
SSK_ERR_CUDA when no GPU is present - there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSK_LIB") or os.path.join(_HERE, "libssk.so")   # SSK_LIB: an alternative build of the same ABI (A/B runs)

if not os.path.exists(LIB_PATH):
    raise ImportError("serstacker_b200/libssk.so is missing: run `python -m serstacker_b200.build` "
                      "(nvcc, sm_100a). There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

SSK_OK, SSK_ERR_INVALID, SSK_ERR_CUDA, SSK_ERR_STATE, SSK_ERR_NOT_REGISTERED, SSK_ERR_NCCL = 0, -1, -2, -3, -4, -5
SSK_8U, SSK_16U, SSK_32F = 0, 2, 5
MEM_HOST, MEM_DEVICE = 0, 1

MOTION_TRANSLATION, MOTION_EUCLIDEAN, MOTION_SCALED_EUCLIDEAN, MOTION_AFFINE, MOTION_HOMOGRAPHY = 0, 1, 2, 3, 4
ECC_FORWARD_ADDITIVE, ECC_INVERSE_COMPOSITIONAL, ECC_LM, ECC_INVERSE_COMPOSITIONAL_LM = 0, 1, 2, 3
ECCFLOW_DOWNSCALE_RECURSIVE_RESIZE, ECCFLOW_DOWNSCALE_FULL_RESIZE, ECCFLOW_DOWNSCALE_PYRAMID = 0, 1, 2
INTER_NEAREST, INTER_LINEAR, INTER_CUBIC, INTER_AREA = 0, 1, 2, 3
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT101, BORDER_TRANSPARENT = 0, 1, 2, 3, 4, 5
ACC_WEIGHTED_AVERAGE, ACC_BAYER_AVERAGE = 0, 1
STACK_AVERAGE, STACK_WEIGHTED_AVERAGE, STACK_BAYER_AVERAGE = 0, 1, 2
UPSCALE_NONE, UPSCALE_PYRUP, UPSCALE_X15, UPSCALE_X30 = 0, 1, 2, 3
UPSCALE_AFTER_ALIGN, UPSCALE_BEFORE_ALIGN = 1, 2
COLORID_MONO = 0
COLORID_BAYER_RGGB, COLORID_BAYER_GRBG, COLORID_BAYER_GBRG, COLORID_BAYER_BGGR = 8, 9, 10, 11


def maketype(depth, cn):
    return (depth & 7) + ((cn - 1) << 3)


class ssk_mat(C.Structure):
    _fields_ = [("data", C.c_void_p), ("step", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32),
                ("type", C.c_int32), ("mem", C.c_int32)]


class ssk_ecch_options(C.Structure):
    _fields_ = [("epsx", C.c_double), ("reference_smooth_sigma", C.c_double), ("input_smooth_sigma", C.c_double),
                ("update_step_scale", C.c_double), ("method", C.c_int32), ("interpolation", C.c_int32),
                ("max_iterations", C.c_int32), ("minimum_image_size", C.c_int32), ("maxlevel", C.c_int32)]


class ssk_ecc_registration_options(C.Structure):
    _fields_ = [("scale", C.c_double), ("eps", C.c_double), ("min_rho", C.c_double),
                ("input_smooth_sigma", C.c_double), ("reference_smooth_sigma", C.c_double),
                ("update_step_scale", C.c_double), ("se_radius", C.c_int32), ("ecc_method", C.c_int32),
                ("max_iterations", C.c_int32), ("ecch_max_level", C.c_int32), ("ecch_minimum_image_size", C.c_int32),
                ("normalization_noise", C.c_double), ("normalization_scale", C.c_int32),
                ("ecch_estimate_translation_first", C.c_int32), ("replace_planetary_disk_with_mask", C.c_int32)]


class ssk_eccflow_options(C.Structure):
    _fields_ = [("input_smooth_sigma", C.c_double), ("reference_smooth_sigma", C.c_double), ("update_multiplier", C.c_double),
                ("scale_factor", C.c_double), ("noise_level", C.c_double), ("max_iterations", C.c_int32),
                ("support_scale", C.c_int32), ("min_image_size", C.c_int32), ("max_pyramid_level", C.c_int32),
                ("downscale_method", C.c_int32), ("reserved", C.c_int32)]


class ssk_registration_options(C.Structure):
    _fields_ = [("motion_type", C.c_int32), ("interpolation", C.c_int32), ("border_mode", C.c_int32),
                ("border_value", C.c_double * 4), ("ecc", ssk_ecc_registration_options),
                ("enable_ecc_registration", C.c_int32), ("enable_eccflow_registration", C.c_int32),
                ("eccflow", ssk_eccflow_options)]


class ssk_ecc_status(C.Structure):
    _fields_ = [("rho", C.c_double), ("min_rho", C.c_double), ("eps", C.c_double), ("num_iterations", C.c_int32),
                ("max_iterations", C.c_int32), ("ok", C.c_int32), ("failed", C.c_int32)]


class ssk_transform(C.Structure):
    _fields_ = [("motion_type", C.c_int32), ("nparams", C.c_int32), ("params", C.c_float * 8), ("aux", C.c_float * 4)]

    def parameters(self):
        return np.array(self.params[:self.nparams], dtype=np.float32)


class ssk_stack_options(C.Structure):
    _fields_ = [("registration", ssk_registration_options), ("accumulation_method", C.c_int32),
                ("sm_dscale", C.c_int32), ("sm_kradius", C.c_int32), ("sm_uscale", C.c_int32),
                ("enable_registration", C.c_int32), ("bayer_colorid", C.c_int32), ("max_batch", C.c_int32),
                ("generating_master_frame", C.c_int32), ("upscale_option", C.c_int32), ("upscale_stage", C.c_int32)]


_P = C.POINTER
_sigs = {
    "ssk_last_error": (C.c_char_p, []),
    "ssk_version": (C.c_int, []),
    "ssk_kernel_launch_count": (C.c_int64, []),
    "ssk_ecch_options_default": (None, [_P(ssk_ecch_options)]),
    "ssk_registration_options_default": (None, [_P(ssk_registration_options)]),
    "ssk_transform_init": (C.c_int, [_P(ssk_transform), C.c_int]),
    "ssk_transform_create_remap": (C.c_int, [_P(ssk_transform), C.c_int, C.c_int, _P(ssk_mat)]),
    "ssk_transform_scale": (C.c_int, [_P(ssk_transform), C.c_double]),
    "ssk_transform_eps": (C.c_int, [_P(ssk_transform), _P(C.c_float), C.c_int, C.c_int, C.c_int, _P(C.c_double)]),
    "ssk_transform_invert_and_compose": (C.c_int, [_P(ssk_transform), _P(C.c_float), C.c_int, _P(C.c_float)]),
    "ssk_transform_remap_points": (C.c_int, [_P(ssk_transform), _P(C.c_float), C.c_int, _P(C.c_float)]),
    "ssk_remap": (C.c_int, [_P(ssk_transform), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat),
                            C.c_int, C.c_int, _P(C.c_double)]),
    "ssk_ecch_create": (C.c_int, [_P(ssk_ecch_options), _P(C.c_void_p)]),
    "ssk_ecch_destroy": (C.c_int, [C.c_void_p]),
    "ssk_ecch_set_reference_image": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat)]),
    "ssk_ecch_align": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), _P(ssk_transform), _P(ssk_ecc_status)]),
    "ssk_ecch_set_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "ssk_ecch_get_trace": (C.c_int, [C.c_void_p, _P(C.c_float), C.c_int, _P(C.c_int)]),
    "ssk_reg_set_trace": (C.c_int, [C.c_void_p, C.c_int]),
    "ssk_reg_get_trace": (C.c_int, [C.c_void_p, _P(C.c_float), C.c_int, _P(C.c_int)]),
    "ssk_ecch_num_levels": (C.c_int, [C.c_void_p]),
    "ssk_ecch_level_size": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int), _P(C.c_int)]),
    "ssk_ecch_get_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(ssk_mat)]),
    "ssk_reg_create": (C.c_int, [_P(ssk_registration_options), _P(C.c_void_p)]),
    "ssk_reg_destroy": (C.c_int, [C.c_void_p]),
    "ssk_reg_setup_reference_frame": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_reg_register_frame": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_int, _P(ssk_transform), _P(ssk_ecc_status)]),
    "ssk_reg_get_current_remap": (C.c_int, [C.c_void_p, _P(ssk_mat)]),
    "ssk_reg_remap": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), C.c_int, C.c_int, _P(C.c_double)]),
    "ssk_acc_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "ssk_acc_destroy": (C.c_int, [C.c_void_p]),
    "ssk_acc_clear": (C.c_int, [C.c_void_p]),
    "ssk_acc_add": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_acc_compute": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_double]),
    "ssk_acc_get_counters": (C.c_int, [C.c_void_p, _P(ssk_mat)]),
    "ssk_acc_reinitialize": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat)]),
    "ssk_acc_size": (C.c_int, [C.c_void_p, _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "ssk_acc_frames": (C.c_int, [C.c_void_p]),
    "ssk_acc_set_bayer_pattern": (C.c_int, [C.c_void_p, C.c_int]),
    "ssk_acc_set_remap": (C.c_int, [C.c_void_p, _P(ssk_transform), _P(ssk_mat)]),
    "ssk_acc_device_state": (C.c_int, [C.c_void_p, _P(C.c_void_p), _P(C.c_void_p), _P(C.c_int64), _P(C.c_int64)]),
    "ssk_acc_to_sum_form": (C.c_int, [C.c_void_p]),
    "ssk_acc_from_sum_form": (C.c_int, [C.c_void_p, C.c_int]),
    "ssk_local_variance_map": (C.c_int, [_P(ssk_mat), C.c_int, C.c_int, C.c_int, C.c_int, _P(ssk_mat), _P(C.c_double)]),
    "ssk_lpg": (C.c_int, [_P(ssk_mat), C.c_double, C.c_double, C.c_int, C.c_int, _P(ssk_mat)]),
    "ssk_gaussian_blur": (C.c_int, [_P(ssk_mat), C.c_double, C.c_double, _P(ssk_mat)]),
    "ssk_debayer_nn2": (C.c_int, [_P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_average_bayer_planes": (C.c_int, [_P(ssk_mat), _P(ssk_mat)]),
    "ssk_input_calibrate": (C.c_int, [_P(ssk_mat), C.c_int, _P(ssk_mat), _P(ssk_mat), _P(ssk_mat)]),
    "ssk_color_transform": (C.c_int, [_P(ssk_mat), _P(C.c_float), C.c_int, _P(ssk_mat)]),
    "ssk_linear_interpolation_inpaint": (C.c_int, [_P(ssk_mat), _P(ssk_mat), _P(ssk_mat)]),
    "ssk_ser_open": (C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    "ssk_ser_close": (C.c_int, [C.c_void_p]),
    "ssk_ser_info": (C.c_int, [C.c_void_p, _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "ssk_ser_read": (C.c_int, [C.c_void_p, C.c_int, _P(ssk_mat), _P(C.c_uint64)]),
    "ssk_unsharp_mask": (C.c_int, [_P(ssk_mat), _P(ssk_mat), C.c_double, C.c_double, C.c_double, C.c_double]),
    "ssk_average_pyramid_inpaint": (C.c_int, [_P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_acc_compute_inpainted": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_double, C.c_int]),
    "ssk_stack_compute_inpainted": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_jdr_derotate_and_add": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), _P(C.c_double), _P(C.c_double), _P(C.c_double),
                                           _P(C.c_double), C.c_double, _P(C.c_int), C.c_double, C.c_int, C.c_int, C.c_double,
                                           C.c_double, C.c_int, C.c_int]),
    "ssk_ellipsoid_zrotation_remap": (C.c_int, [C.c_int, C.c_int, _P(C.c_double), _P(C.c_double), _P(C.c_double), _P(C.c_double),
                                                C.c_double, _P(C.c_int), C.c_double, _P(ssk_mat), _P(ssk_mat), _P(ssk_mat)]),
    "ssk_stack_options_default": (None, [_P(ssk_stack_options)]),
    "ssk_stack_create": (C.c_int, [_P(ssk_stack_options), _P(C.c_void_p)]),
    "ssk_stack_destroy": (C.c_int, [C.c_void_p]),
    "ssk_stack_set_reference": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_stack_add_frames": (C.c_int, [C.c_void_p, _P(ssk_mat), C.c_int, C.c_int, _P(ssk_transform), _P(ssk_ecc_status)]),
    "ssk_stack_add_frames_async": (C.c_int, [C.c_void_p, _P(ssk_mat), C.c_int, C.c_int]),
    "ssk_stack_sync": (C.c_int, [C.c_void_p]),
    "ssk_stack_flush": (C.c_int, [C.c_void_p]),
    "ssk_stack_submit": (C.c_int, [C.c_void_p, _P(ssk_mat), C.c_int, C.c_int, _P(C.c_int64)]),
    "ssk_stack_wait": (C.c_int, [C.c_void_p, C.c_int64, _P(ssk_transform), _P(ssk_ecc_status), C.c_int, _P(C.c_int)]),
    "ssk_stack_compute": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat)]),
    "ssk_stack_accumulated_frames": (C.c_int, [C.c_void_p]),
    "ssk_stack_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ssk_stack_reset": (C.c_int, [C.c_void_p]),
    "ssk_acc_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ssk_nccl_get_unique_id": (C.c_int, [C.c_void_p]),
    "ssk_nccl_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p)]),
    "ssk_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "ssk_stack_accumulator": (C.c_void_p, [C.c_void_p]),
    "ssk_stack_registration": (C.c_void_p, [C.c_void_p]),
    "ssk_stack_stream": (C.c_void_p, [C.c_void_p]),
    "ssk_stack_stage_times": (C.c_int, [C.c_void_p, _P(C.c_float)]),
    "ssk_median_filter_bad_pixels": (C.c_int, [_P(ssk_mat), C.c_double]),
    "ssk_bayer_denoise": (C.c_int, [_P(ssk_mat), C.c_double]),
    "ssk_set_stream_ordered": (C.c_int, [C.c_int]),
    "ssk_device_synchronize": (C.c_int, []),
    "ssk_build_ellipsoid_rotation": (C.c_int, [_P(C.c_double), _P(C.c_double)]),
    "ssk_ellipsoid_bbox": (C.c_int, [C.c_int, C.c_int, _P(C.c_double), _P(C.c_double), _P(C.c_double), _P(C.c_float), _P(C.c_int)]),
    "ssk_upscale_size": (C.c_int, [C.c_int, C.c_int, C.c_int, _P(C.c_int), _P(C.c_int)]),
    "ssk_upscale_image": (C.c_int, [C.c_int, _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(ssk_mat)]),
    "ssk_upscale_remap": (C.c_int, [C.c_int, _P(ssk_mat), _P(ssk_mat)]),
    "ssk_upscale_optflow": (C.c_int, [C.c_int, _P(ssk_mat), _P(ssk_mat)]),
    "ssk_canvas_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "ssk_canvas_destroy": (C.c_int, [C.c_void_p]),
    "ssk_canvas_set_canvas_size": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ssk_canvas_add": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), _P(C.c_int)]),
    "ssk_canvas_compute": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), C.c_double, _P(C.c_int)]),
    "ssk_canvas_clear": (C.c_int, [C.c_void_p]),
    "ssk_canvas_accumulated_frames": (C.c_int, [C.c_void_p]),
    "ssk_canvas_size": (C.c_int, [C.c_void_p, _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "ssk_canvas_last_bbox": (C.c_int, [C.c_void_p, _P(C.c_int)]),
    "ssk_eccflow_options_default": (None, [_P(ssk_eccflow_options)]),
    "ssk_eccflow_registration_options_default": (None, [_P(ssk_eccflow_options)]),
    "ssk_eccflow_create": (C.c_int, [_P(ssk_eccflow_options), _P(C.c_void_p)]),
    "ssk_eccflow_destroy": (C.c_int, [C.c_void_p]),
    "ssk_eccflow_set_reference_image": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat)]),
    "ssk_eccflow_compute": (C.c_int, [C.c_void_p, _P(ssk_mat), _P(ssk_mat), _P(ssk_mat), C.c_int]),
    "ssk_eccflow_get_uv": (C.c_int, [C.c_void_p, _P(ssk_mat)]),
    "ssk_eccflow_num_levels": (C.c_int, [C.c_void_p]),
    "ssk_eccflow_level_size": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_int)]),
    "ssk_eccflow_get_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(ssk_mat)]),
}
for _name, (_res, _args) in _sigs.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args

EXPORTED = sorted(_sigs)


class SskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ssk error %d: %s" % (code, msg))
        self.code = code


def last_error():
    return (lib.ssk_last_error() or b"").decode("utf-8", "replace")


def check(code):
    if code != SSK_OK:
        raise SskError(code, last_error())
    return code


_DEPTH = {np.dtype(np.uint8): SSK_8U, np.dtype(np.uint16): SSK_16U, np.dtype(np.float32): SSK_32F}


def mat(a):
    """ssk_mat view of a numpy array (HxW or HxWxC; rows may be strided, pixels must be packed)."""
    if a is None:
        return None
    assert a.dtype in _DEPTH, a.dtype
    cn = 1 if a.ndim == 2 else a.shape[2]
    es = a.dtype.itemsize
    assert a.strides[-1] == es and (a.ndim == 2 or a.strides[1] == es * cn), "pixels must be packed"
    m = ssk_mat(a.ctypes.data, a.strides[0], a.shape[0], a.shape[1], maketype(_DEPTH[a.dtype], cn), MEM_HOST)
    m._keep = a
    return m


def device_mat(ptr, rows, cols, dtype, cn=1, step=None):
    """ssk_mat view of device memory (e.g. torch_tensor.data_ptr())."""
    dt = np.dtype(dtype)
    return ssk_mat(ptr, step if step is not None else cols * cn * dt.itemsize, rows, cols, maketype(_DEPTH[dt], cn), MEM_DEVICE)


def ref(m):
    return None if m is None else C.byref(m)
