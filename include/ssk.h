/*
 * ssk.h - C ABI of the B200-native SerStacker stacking hot path (register -> warp -> accumulate).
 *
 * The reference (amyznikov/SerStacker) has no FFI layer: the boundary of this path is the C++ class
 * surface its pipelines call.  Each entry point below names the reference member it replaces
 * (paths relative to the reference root).  A thin C++ adapter with the reference's class names lives in
 * serstacker_b200/host/ssk_adapter.h and forwards to these functions; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns 0 (SSK_OK) or a negative ssk_status and never
 *    throws; ssk_last_error() returns the message of the last failure on the calling thread
 *    (reference: bool return + CF_ERROR log line, core/debug.h:88-95).
 *  - images are described by ssk_mat: the memory layout of cv::Mat (row-major, `step` bytes per row,
 *    interleaved channels, OpenCV type code).  `mem` says whether `data` is a host or a device pointer.
 *    Inputs are borrowed for the duration of the call; outputs are written into caller-allocated buffers.
 *  - handles are single-threaded (one CUDA stream per handle), like the reference objects.
 *  - there is no CPU fallback: every call fails with SSK_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef SSK_H_
#define SSK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSK_API __attribute__((visibility("default")))

typedef enum ssk_status {
  SSK_OK = 0,
  SSK_ERR_INVALID = -1,      /* bad argument / unsupported combination */
  SSK_ERR_CUDA = -2,         /* CUDA runtime failure (message in ssk_last_error) */
  SSK_ERR_STATE = -3,        /* call order violated (e.g. no reference set) */
  SSK_ERR_NOT_REGISTERED = -4, /* frame failed registration (low rho / singular H): reference returns false */
  SSK_ERR_NCCL = -5
} ssk_status;

/* OpenCV type codes used by the path (CV_MAKETYPE(depth, cn)). */
enum { SSK_8U = 0, SSK_16U = 2, SSK_32F = 5 };
#define SSK_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << 3))
#define SSK_8UC1 SSK_MAKETYPE(SSK_8U, 1)
#define SSK_16UC1 SSK_MAKETYPE(SSK_16U, 1)
#define SSK_32FC1 SSK_MAKETYPE(SSK_32F, 1)
#define SSK_32FC2 SSK_MAKETYPE(SSK_32F, 2)
#define SSK_32FC3 SSK_MAKETYPE(SSK_32F, 3)
#define SSK_32FC4 SSK_MAKETYPE(SSK_32F, 4)

enum { SSK_MEM_HOST = 0, SSK_MEM_DEVICE = 1 };

/* cv::Mat view. */
typedef struct ssk_mat {
  void *data;
  int64_t step;   /* bytes per row */
  int32_t rows, cols;
  int32_t type;   /* OpenCV type code */
  int32_t mem;    /* SSK_MEM_HOST / SSK_MEM_DEVICE */
} ssk_mat;

/* IMAGE_MOTION_TYPE, core/proc/image_registration/image_transform.h:14-25 (same values). */
enum {
  SSK_MOTION_TRANSLATION = 0,
  SSK_MOTION_EUCLIDEAN = 1,
  SSK_MOTION_SCALED_EUCLIDEAN = 2,
  SSK_MOTION_AFFINE = 3,
  SSK_MOTION_HOMOGRAPHY = 4
};

/* ECC_ALIGN_METHOD, ecc2.h:59-64 (same values). */
enum {
  SSK_ECC_FORWARD_ADDITIVE = 0,
  SSK_ECC_INVERSE_COMPOSITIONAL = 1,
  SSK_ECC_LM = 2,
  SSK_ECC_INVERSE_COMPOSITIONAL_LM = 3
};

/* ECC_INTERPOLATION_METHOD / ECC_BORDER_MODE, ecc2.h:27-52 (cv::InterpolationFlags / cv::BorderTypes values). */
/* cv::InterpolationFlags values.  cv::remap itself replaces INTER_AREA by INTER_LINEAR, so SSK_INTER_AREA is accepted
   wherever a remap interpolation is expected and behaves as SSK_INTER_LINEAR (ECC_INTER_AREA, ecc2.h:37). */
enum { SSK_INTER_NEAREST = 0, SSK_INTER_LINEAR = 1, SSK_INTER_CUBIC = 2, SSK_INTER_AREA = 3, SSK_INTER_LANCZOS4 = 4 };
enum {
  SSK_BORDER_CONSTANT = 0, SSK_BORDER_REPLICATE = 1, SSK_BORDER_REFLECT = 2, SSK_BORDER_WRAP = 3,
  SSK_BORDER_REFLECT101 = 4, SSK_BORDER_TRANSPARENT = 5
};

/* COLORID, core/io/debayer.h:19-35 (same values). */
enum { SSK_COLORID_MONO = 0, SSK_COLORID_BAYER_RGGB = 8, SSK_COLORID_BAYER_GRBG = 9,
       SSK_COLORID_BAYER_GBRG = 10, SSK_COLORID_BAYER_BGGR = 11 };

#define SSK_MAX_PARAMS 8

/* c_ecch_options, ecc2.h:158-172 (same fields, same defaults via ssk_ecch_options_default). */
typedef struct ssk_ecch_options {
  double epsx;
  double reference_smooth_sigma;
  double input_smooth_sigma;
  double update_step_scale;
  int32_t method;
  int32_t interpolation;
  int32_t max_iterations;
  int32_t minimum_image_size;
  int32_t maxlevel;
} ssk_ecch_options;

/* c_ecc_registration_options, c_frame_registration.h:47-64. */
typedef struct ssk_ecc_registration_options {
  double scale;
  double eps;
  double min_rho;
  double input_smooth_sigma;
  double reference_smooth_sigma;
  double update_step_scale;
  int32_t se_radius;
  int32_t ecc_method;
  int32_t max_iterations;
  int32_t ecch_max_level;
  int32_t ecch_minimum_image_size;
  double normalization_noise;
  int32_t normalization_scale;
  int32_t ecch_estimate_translation_first;
  int32_t replace_planetary_disk_with_mask;
} ssk_ecc_registration_options;

/* ECCFlowDownscaleMethod, ecc2.h:503-507 (same order). */
enum { SSK_ECCFLOW_DOWNSCALE_RECURSIVE_RESIZE = 0, SSK_ECCFLOW_DOWNSCALE_FULL_RESIZE = 1, SSK_ECCFLOW_DOWNSCALE_PYRAMID = 2 };

/* c_eccflow_options, ecc2.h:515-527 (defaults via ssk_eccflow_options_default), which is also the layout of
 * c_eccflow_registration_options, c_frame_registration.h:88-100 (defaults via ssk_eccflow_registration_options_default).
 * input_smooth_sigma / reference_smooth_sigma are carried but, as in the reference, not used by the computation. */
typedef struct ssk_eccflow_options {
  double input_smooth_sigma;
  double reference_smooth_sigma;
  double update_multiplier;
  double scale_factor;
  double noise_level;
  int32_t max_iterations;
  int32_t support_scale;
  int32_t min_image_size;
  int32_t max_pyramid_level;
  int32_t downscale_method;
  int32_t reserved;
} ssk_eccflow_options;

/* c_image_registration_options, c_frame_registration.h:119-136 (ECC and eccflow members; the sparse-feature stage is out of
 * scope and must stay disabled). */
typedef struct ssk_registration_options {
  int32_t motion_type;
  int32_t interpolation;
  int32_t border_mode;
  double border_value[4];
  ssk_ecc_registration_options ecc;
  int32_t enable_ecc_registration;
  int32_t enable_eccflow_registration;   /* c_frame_registration.cc:900-917: refine _current_remap per pixel after ECC */
  ssk_eccflow_options eccflow;
} ssk_registration_options;

/* c_image_registration_status::ecc, c_frame_registration.h:198-207. */
typedef struct ssk_ecc_status {
  double rho;
  double min_rho;
  double eps;
  int32_t num_iterations;
  int32_t max_iterations;
  int32_t ok;       /* 1: registered; 0: dropped (reference: register_frame() returned false) */
  int32_t failed;   /* solver flagged failure (singular Hessian), c_ecc_align::failed() */
} ssk_ecc_status;

/* An image transform by value: what c_image_transform holds (c_image_transform.h:42-117).
 * params layout = the reference's `parameters()` vector:
 *   translation (tx,ty); euclidean (tx,ty,angle[,scale]); affine a00 a01 a02 a10 a11 a12;
 *   homography a00..a21 (a22 kept in aux[2]).  aux = {Cx, Cy, a22, euclidean scale when fixed}. */
typedef struct ssk_transform {
  int32_t motion_type;
  int32_t nparams;
  float params[SSK_MAX_PARAMS];
  float aux[4];
} ssk_transform;

SSK_API const char *ssk_last_error(void);
SSK_API int ssk_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
SSK_API int64_t ssk_kernel_launch_count(void);

SSK_API void ssk_ecch_options_default(ssk_ecch_options *o);                  /* ecc2.h:158-172 */
SSK_API void ssk_registration_options_default(ssk_registration_options *o);  /* c_frame_registration.h:47-64,119-136 */
SSK_API int ssk_transform_init(ssk_transform *t, int motion_type);           /* image_transform.cc:34-63 + reset() */

/* ---------------------------------------------------------------------------------------------
 * c_image_transform (stateless helpers; c_image_transform.cc)
 * ------------------------------------------------------------------------------------------- */
/* c_image_transform::create_remap(params, size, rmap): rmap is CV_32FC2 rows x cols. */
SSK_API int ssk_transform_create_remap(const ssk_transform *t, int rows, int cols, ssk_mat *rmap);
/* c_image_transform::scale_transfrom(factor). */
SSK_API int ssk_transform_scale(ssk_transform *t, double factor);
/* c_image_transform::eps(dp, image_size): the size of a parameter step in pixels, as the solvers' convergence test reads it
 * (c_image_transform.cc:136-139, 509-523, 917-924, 1196-1205).  Host arithmetic, no device call; ndp = t->nparams. */
SSK_API int ssk_transform_eps(const ssk_transform *t, const float *dp, int ndp, int rows, int cols, double *eps);
/* c_image_transform::invert_and_compose(parameters(), dp): the parameters of  W(p) o W(dp)^-1  (c_image_transform.h:164-167,
 * 312-318, 379-384; c_image_transform.cc:736-833); out receives t->nparams floats (homography: a22 normalised to 1). Host arithmetic. */
SSK_API int ssk_transform_invert_and_compose(const ssk_transform *t, const float *dp, int ndp, float *out);
/* c_image_transform::remap(parameters(), rpts, cpts): n reference points (x, y interleaved) -> their positions in the current frame
 * (c_image_transform.cc:232-249, 557-585, 1019-1033, 1294-1306).  Host arithmetic. */
SSK_API int ssk_transform_remap_points(const ssk_transform *t, const float *rpts_xy, int n, float *cpts_xy);

/* ---------------------------------------------------------------------------------------------
 * cv::remap as the reference calls it (c_frame_registration::base_remap, c_frame_registration.cc:1265-1386):
 * dst = remap(src); if dst_mask != NULL: dst_mask = erode5x5(remap(src_mask or all-255, interp, CONSTANT 0) >= 255).
 * The map is either analytic (t != NULL) or explicit (rmap != NULL, CV_32FC2).
 * ------------------------------------------------------------------------------------------- */
SSK_API int ssk_remap(const ssk_transform *t, const ssk_mat *rmap,
                      const ssk_mat *src, ssk_mat *dst,
                      const ssk_mat *src_mask, ssk_mat *dst_mask,
                      int interpolation, int border_mode, const double border_value[4]);

/* ---------------------------------------------------------------------------------------------
 * c_ecch (ecc2.h:193-308, ecc2.cc:695-1176): coarse-to-fine ECC against a fixed reference.
 * ------------------------------------------------------------------------------------------- */
typedef struct ssk_ecch ssk_ecch;
SSK_API int ssk_ecch_create(const ssk_ecch_options *opts, ssk_ecch **out);
SSK_API int ssk_ecch_destroy(ssk_ecch *h);
/* c_ecch::set_reference_image(image, mask): image CV_32FC1 (or 8U/16U, converted), mask CV_8UC1 or NULL. */
SSK_API int ssk_ecch_set_reference_image(ssk_ecch *h, const ssk_mat *image, const ssk_mat *mask);
/* c_ecch::align(current_image, current_mask): t is the c_image_transform the reference binds with
 * set_image_transform(); it is updated in place.  status may be NULL. */
SSK_API int ssk_ecch_align(ssk_ecch *h, const ssk_mat *image, const ssk_mat *mask, ssk_transform *t,
                           ssk_ecc_status *status);
/* Debug aid (the reference dumps registration artefacts under <out>/debug, c_image_stacking_pipeline.cc:1511-1530):
 * record every solver trial of the next align() calls: SSK_TRACE_REC floats per record =
 * level, pass, num_it, err, newerr, lambda, eps, n_valid | accepted params[8] | trial params[8] | deltap[8] | v[8]. */
#define SSK_TRACE_REC 40
SSK_API int ssk_ecch_set_trace(ssk_ecch *h, int max_records);
SSK_API int ssk_ecch_get_trace(ssk_ecch *h, float *records, int max_records, int *n);
/* number of pyramid levels / size of level l (c_ecch::compute_next_pyramid_layer_size, ecc2.h:290-293). */
SSK_API int ssk_ecch_num_levels(const ssk_ecch *h);
SSK_API int ssk_ecch_level_size(const ssk_ecch *h, int level, int *cols, int *rows);
/* c_ecch::reference_image()/current_image(): copies pyramid level `level` (CV_32FC1) into dst. */
SSK_API int ssk_ecch_get_image(const ssk_ecch *h, int which /*0 reference, 1 current*/, int level, ssk_mat *dst);

/* ---------------------------------------------------------------------------------------------
 * c_frame_registration, ECC branch (c_frame_registration.h:210-354, c_frame_registration.cc:565-964).
 * ------------------------------------------------------------------------------------------- */
typedef struct ssk_reg ssk_reg;
SSK_API int ssk_reg_create(const ssk_registration_options *opts, ssk_reg **out);
SSK_API int ssk_reg_destroy(ssk_reg *h);
/* c_frame_registration::setup_reference_frame(image, mask). Frames are what the pipeline feeds:
 * CV_32F in [0,1) (or 8U/16U with `bpp`, normalised by 1/(1<<bpp) as c_image_stacking_pipeline_base.cc:271-276). */
SSK_API int ssk_reg_setup_reference_frame(ssk_reg *h, const ssk_mat *image, const ssk_mat *mask, int bpp);
/* c_frame_registration::register_frame(src, srcmask) without dst: estimates the transform only.
 * Returns SSK_ERR_NOT_REGISTERED where the reference returns false. */
SSK_API int ssk_reg_register_frame(ssk_reg *h, const ssk_mat *image, const ssk_mat *mask, int bpp,
                                   ssk_transform *t_out, ssk_ecc_status *status);
SSK_API int ssk_reg_set_trace(ssk_reg *h, int max_records);
SSK_API int ssk_reg_get_trace(ssk_reg *h, float *records, int max_records, int *n);
/* c_frame_registration::current_remap(): materialises the full-resolution CV_32FC2 map on request. */
SSK_API int ssk_reg_get_current_remap(ssk_reg *h, ssk_mat *rmap);
/* c_frame_registration::remap()/custom_remap() with the current transform (rmap==NULL) or an explicit map. */
SSK_API int ssk_reg_remap(ssk_reg *h, const ssk_mat *rmap, const ssk_mat *src, ssk_mat *dst,
                          const ssk_mat *src_mask, ssk_mat *dst_mask,
                          int interpolation /*<0: options*/, int border_mode /*<0: options*/,
                          const double border_value[4]);

/* ---------------------------------------------------------------------------------------------
 * c_image_stacking_pipeline::upscale_image / upscale_remap / upscale_optflow (c_image_stacking_pipeline.cc:1869-2002):
 * option = SSK_UPSCALE_* (x2.0 = cv::pyrUp, x1.5 = cv::resize INTER_LINEAR, x3.0 = cv::resize INTER_LINEAR_EXACT).
 * ------------------------------------------------------------------------------------------- */
SSK_API int ssk_upscale_size(int option, int cols, int rows, int *ucols, int *urows);
/* src / dst CV_32F (1-4 channels), srcmask / dstmask CV_8UC1 (result compared >= 255 as the reference does); either pair may be NULL. */
SSK_API int ssk_upscale_image(int option, const ssk_mat *src, const ssk_mat *srcmask, ssk_mat *dst, ssk_mat *dstmask);
SSK_API int ssk_upscale_remap(int option, const ssk_mat *srcmap, ssk_mat *dstmap);      /* CV_32FC2 */
SSK_API int ssk_upscale_optflow(int option, const ssk_mat *srcmap, ssk_mat *dstmap);    /* CV_32FC2, values scaled by the factor */

/* ---------------------------------------------------------------------------------------------
 * c_canvas_average (core/average/c_frame_accumulation.h:65-137, c_frame_accumulation.cc:264-445): weighted average on a
 * canvas larger than the frames; used by c_canvas_average_pipeline.  Frames are CV_32F (1-4 channels).
 * ------------------------------------------------------------------------------------------- */
typedef struct ssk_canvas ssk_canvas;
/* interpolation = c_canvas_average::options::interpolation (cv::InterpolationFlags, default INTER_LINEAR). */
SSK_API int ssk_canvas_create(int interpolation, ssk_canvas **out);
SSK_API int ssk_canvas_destroy(ssk_canvas *h);
SSK_API int ssk_canvas_set_canvas_size(ssk_canvas *h, int cols, int rows);   /* setCanvasSize(): clears */
/* add(current_image, current_weights_or_mask, rmap, new_canvas_bbox): weights NULL / CV_8UC1 mask / CV_32FC1 weights of the image
 * size; rmap NULL or CV_32FC2 of the box size; bbox = {x, y, width, height} or NULL.  The first frame is centred on a new canvas
 * of max(setCanvasSize, 3/2 frame size). */
SSK_API int ssk_canvas_add(ssk_canvas *h, const ssk_mat *image, const ssk_mat *weights_or_mask, const ssk_mat *rmap, const int bbox[4]);
/* compute(avg, mask, dscale, ddepth = CV_32F, rbbox): the box (NULL: whole canvas) clipped to the canvas; avg / mask must have
 * that size (ssk_canvas_size gives the canvas size). */
SSK_API int ssk_canvas_compute(ssk_canvas *h, ssk_mat *avg, ssk_mat *mask, double dscale, const int rbbox[4]);
SSK_API int ssk_canvas_clear(ssk_canvas *h);
SSK_API int ssk_canvas_accumulated_frames(const ssk_canvas *h);
SSK_API int ssk_canvas_size(const ssk_canvas *h, int *cols, int *rows, int *channels);   /* accumulator_size() */
SSK_API int ssk_canvas_last_bbox(const ssk_canvas *h, int bbox[4]);                       /* last_bbox() */

/* ---------------------------------------------------------------------------------------------
 * c_eccflow (ecc2.h:548-662, ecc2.cc:2220-2865): dense smooth optical flow on a coarse-to-fine pyramid.
 * ------------------------------------------------------------------------------------------- */
typedef struct ssk_eccflow ssk_eccflow;
SSK_API void ssk_eccflow_options_default(ssk_eccflow_options *o);                /* ecc2.h:515-527 */
SSK_API void ssk_eccflow_registration_options_default(ssk_eccflow_options *o);   /* c_frame_registration.h:88-100 */
SSK_API int ssk_eccflow_create(const ssk_eccflow_options *opts, ssk_eccflow **out);
SSK_API int ssk_eccflow_destroy(ssk_eccflow *h);
/* c_eccflow::set_reference_image(reference_image, reference_mask): single-channel image (any depth, converted to CV_32F
 * without scaling as convertTo does), optional CV_8UC1 mask. */
SSK_API int ssk_eccflow_set_reference_image(ssk_eccflow *h, const ssk_mat *image, const ssk_mat *mask);
/* c_eccflow::compute(input_image, rmap, input_mask): rmap is CV_32FC2 of the reference size, read as the initial map when
 * use_initial_map != 0 (an "empty rmap" otherwise) and overwritten with the refined map. */
SSK_API int ssk_eccflow_compute(ssk_eccflow *h, const ssk_mat *image, const ssk_mat *mask, ssk_mat *rmap, int use_initial_map);
/* c_eccflow::current_uv(): the flow of the last compute (CV_32FC2 of the reference size). */
SSK_API int ssk_eccflow_get_uv(ssk_eccflow *h, ssk_mat *uv);
/* c_eccflow::current_pyramid() (debug / tests): number of levels, level geometry (size of the level and of its avgdown grid),
 * level images: which = 0 reference_image, 1 current_image, 2 Ix, 3 Iy (CV_32FC1 of the level size), 4 D (CV_32FC4, grid size). */
SSK_API int ssk_eccflow_num_levels(const ssk_eccflow *h);
SSK_API int ssk_eccflow_level_size(const ssk_eccflow *h, int level, int *cols, int *rows, int *grid_cols, int *grid_rows);
SSK_API int ssk_eccflow_get_image(ssk_eccflow *h, int which, int level, ssk_mat *dst);

/* ---------------------------------------------------------------------------------------------
 * c_frame_accumulation (core/average/c_frame_accumulation.h:14-63, 222-262).
 * ------------------------------------------------------------------------------------------- */
typedef struct ssk_acc ssk_acc;
enum { SSK_ACC_WEIGHTED_AVERAGE = 0,   /* c_weigthed_average (also plain `average`) */
       SSK_ACC_BAYER_AVERAGE = 1 };    /* c_bayer_average */
SSK_API int ssk_acc_create(int kind, ssk_acc **out);
SSK_API int ssk_acc_destroy(ssk_acc *h);
SSK_API int ssk_acc_clear(ssk_acc *h);                                       /* clear() */
/* add(src, mask_or_weights): weights NULL | CV_8UC1 mask | CV_32FC1 weights (c_frame_accumulation.cc:20-129). */
SSK_API int ssk_acc_add(ssk_acc *h, const ssk_mat *src, const ssk_mat *weights, int bpp);
/* compute(avg, mask, dscale, ddepth): avg CV_32F (cn channels), mask CV_8UC1 (NULL to skip). */
SSK_API int ssk_acc_compute(ssk_acc *h, ssk_mat *avg, ssk_mat *mask, double dscale);
SSK_API int ssk_acc_get_counters(ssk_acc *h, ssk_mat *accw);                  /* get_acc_counters() */
SSK_API int ssk_acc_reinitialize(ssk_acc *h, const ssk_mat *src, const ssk_mat *accw); /* reinitialize() */
SSK_API int ssk_acc_size(const ssk_acc *h, int *cols, int *rows, int *channels); /* accumulator_size() */
SSK_API int ssk_acc_frames(const ssk_acc *h);                                 /* accumulated_frames() */
/* c_bayer_average::set_bayer_pattern / set_remap (remap by transform or explicit CV_32FC2 map). */
SSK_API int ssk_acc_set_bayer_pattern(ssk_acc *h, int colorid);
SSK_API int ssk_acc_set_remap(ssk_acc *h, const ssk_transform *t, const ssk_mat *rmap);
/* Device pointers of the state, for the multi-GPU combine: mean (or sum) planes and weight/counter planes. */
SSK_API int ssk_acc_device_state(ssk_acc *h, void **acc, void **weights, int64_t *acc_bytes, int64_t *weights_bytes);
/* Multi-GPU: convert the local running mean to sum form (A*W, W) in place / back (after an external
 * reduce of both buffers, e.g. ncclReduce), so that A = sum(W_g A_g) / sum(W_g). */
SSK_API int ssk_acc_to_sum_form(ssk_acc *h);
SSK_API int ssk_acc_from_sum_form(ssk_acc *h, int accumulated_frames);

/* Multi-GPU combine (frames sharded over ranks, one process per GPU; the reference is single-process and ends a run with
 * ONE c_frame_accumulation, c_frame_accumulation.h:14-63).  nccl_comm is the caller's ncclComm_t.  On return the root's
 * accumulator holds the stack of all ranks' frames (A = sum_g W_g A_g / sum_g W_g, W = sum_g W_g; Bayer sums and counters
 * added) and accumulated_frames() the total; the other ranks keep their local state.  One ncclReduce group (values,
 * weights, frame count) on the handle's stream; NCCL is loaded at run time (libnccl.so.2), SSK_ERR_NCCL if absent. */
SSK_API int ssk_acc_reduce(ssk_acc *h, void *nccl_comm, int root);
/* Conveniences for hosts that do not link NCCL themselves: ncclGetUniqueId (128 bytes), ncclCommInitRank on the current
 * device, ncclCommDestroy. */
SSK_API int ssk_nccl_get_unique_id(void *id128);
SSK_API int ssk_nccl_comm_create(const void *id128, int nranks, int rank, void **nccl_comm);
SSK_API int ssk_nccl_comm_destroy(void *nccl_comm);

/* ---------------------------------------------------------------------------------------------
 * Weight maps (core/proc/sharpness_measure/c_local_variance_sharpness_measure.cc:193-247, core/proc/lpg.cc:223-290).
 * ------------------------------------------------------------------------------------------- */
/* compute_local_variance_map(image, dscale, kradius, uscale) -> sharpness metric *Q and (map != NULL) the weight map.
 * bpp >= 0: integer frames are normalised by 1 / (1 << bpp) first, as read_input_frame hands them to compute_weights;
 * bpp < 0: by 1 / maxval(depth), the reference's own scaling of integer images (select_master_frame ranks raw frames). */
SSK_API int ssk_local_variance_map(const ssk_mat *image, int bpp, int dscale, int kradius, int uscale,
                                   ssk_mat *map /*CV_32FC1, full resolution, or NULL*/, double *Q);
/* lpg(image, k, p, dscale, uscale, map) (core/proc/lpg.cc:223-290; callers c_jdr_pipeline.cc:1211, c_sdr_pipeline.cc:1205):
 * Laplacian + gradient energy weight map, CV_32FC1 of the image size.  Integer powers p only. */
SSK_API int ssk_lpg(const ssk_mat *image, double k, double p, int dscale, int uscale, ssk_mat *map /*CV_32FC1*/);
/* cv::GaussianBlur(src, dst, Size(), sigma_x, sigma_y, BORDER_REPLICATE) on CV_32FC1: the smoothing of the per-frame
 * weights in c_jdr_pipeline::derotate_and_average_frames (c_jdr_pipeline.cc:1228). sigma_y <= 0: sigma_x. */
SSK_API int ssk_gaussian_blur(const ssk_mat *src, double sigma_x, double sigma_y, ssk_mat *dst);

/* ---------------------------------------------------------------------------------------------
 * Input side: debayer_nn2(src, dst, colorid) (core/io/debayer.cc:827-1195), the bilinear demosaic of raw Bayer frames in
 * read_input_frame (c_image_stacking_pipeline_base.cc:125-279).  src: CV_8UC1 / CV_16UC1 / CV_32FC1 with even size;
 * dst: 3 channels (BGR) of the same depth; colorid: SSK_COLORID_BAYER_{RGGB,GRBG,GBRG,BGGR}.
 * ------------------------------------------------------------------------------------------- */
SSK_API int ssk_debayer_nn2(const ssk_mat *src, ssk_mat *dst, int colorid);

/* average_bayer_planes(src, dst) (core/io/debayer.cc:277-376), raw single-channel form: one sample per 2x2 Bayer cell,
 * (2 + s00 + s01 + s10 + s11) / 4 in integer arithmetic (float: the plain mean).  The gray proxy select_master_frame ranks
 * for raw Bayer sequences (c_image_stacking_pipeline_base.cc:370-378).  dst: half the size, same type. */
SSK_API int ssk_average_bayer_planes(const ssk_mat *src, ssk_mat *dst);
/* read_input_frame's dark / flat correction (c_image_stacking_pipeline_base.cc:143-184): dst = (float(frame) / (1 << bpp) - dark)
 * / flat, either of dark / flat may be NULL; dark, flat and dst are CV_32F of the frame's size and channel count. */
SSK_API int ssk_input_calibrate(const ssk_mat *frame, int bpp, const ssk_mat *dark, const ssk_mat *flat, ssk_mat *dst);
/* cv::transform(image, image, color_matrix) of read_input_frame (c_image_stacking_pipeline_base.cc:263-266): CV_32FC3 image,
 * row-major 3 x mcols float matrix (mcols 3 or 4). */
SSK_API int ssk_color_transform(const ssk_mat *src, const float *m, int mcols, ssk_mat *dst);
/* linear_interpolation_inpaint(src, mask, dst) (core/proc/inpaint/linear_interpolation_inpaint.cc:327-368): holes (mask == 0)
 * are filled by the distance-weighted mix of the linear interpolations along their row and their column; applied to the
 * generated master frame (c_image_stacking_pipeline.cc:1282-1284) and to input frames with a missing-pixel mask
 * (c_image_stacking_pipeline_base.cc:258-261).  CV_32F, 1 to 4 channels; mask CV_8UC1 or NULL (plain copy). */
SSK_API int ssk_linear_interpolation_inpaint(const ssk_mat *src, const ssk_mat *mask, ssk_mat *dst);
/* median_filter_bad_pixels(image, variation_threshold, COLORID_MONO / a colour image) (core/proc/bad_pixels.cc:14-70; read_input_frame's
 * filter_bad_pixels option, c_image_stacking_pipeline_base.cc:192-212): in place; 8U / 16U / 32F, 1-4 interleaved channels.
 * A raw Bayer frame goes through ssk_bayer_denoise, as median_filter_bad_pixels does for a Bayer COLORID (bad_pixels.cc:62-64). */
SSK_API int ssk_median_filter_bad_pixels(ssk_mat *image, double variation_threshold);
/* bayer_denoise(image, variation_threshold, colorid, returnBayerPlanes = false) (core/io/debayer.cc:1471-1611; the Bayer branch of
 * filter_bad_pixels, c_image_stacking_pipeline_base.cc:204-209): 3 x 3 median / mean-absolute-deviation test on each of the four
 * colour planes of the raw mosaic, in place; single channel 8U / 16U / 32F, even size.  The CFA order does not enter the arithmetic. */
SSK_API int ssk_bayer_denoise(ssk_mat *image, double variation_threshold);

/* ---------------------------------------------------------------------------------------------
 * SER container (c_ser_reader, core/io/c_ser_file.cc:272-531; 178-byte header c_ser_file.h:42-56, frames back to back,
 * optional trailer of uint64 time stamps; the stored endianness flag is inverted, c_ser_file.cc:305).  Host side.
 * type: OpenCV type code of a frame (CV_8U / CV_16U / CV_32F for bits_per_plane 1..8 / 9..16 / -32; 3 channels for RGB / BGR).
 * ------------------------------------------------------------------------------------------- */
typedef struct ssk_ser ssk_ser;
SSK_API int ssk_ser_open(const char *path, ssk_ser **out);
SSK_API int ssk_ser_close(ssk_ser *s);
SSK_API int ssk_ser_info(const ssk_ser *s, int *cols, int *rows, int *type, int *bits_per_plane, int *color_id, int *frames,
                         int *has_timestamps);
/* c_ser_reader::seek + read: frame `frame_index` into a host image of the file's size and type (pinned memory lets
 * ssk_stack_submit upload it asynchronously); *timestamp: the frame's trailer entry (0 without a trailer), may be NULL. */
SSK_API int ssk_ser_read(ssk_ser *s, int frame_index, ssk_mat *dst, uint64_t *timestamp);

/* ---------------------------------------------------------------------------------------------
 * unsharp_mask(src, dst, sigma, alpha, outmin, outmax) (core/proc/unsharp_mask.cc:72-118): the sharpening applied to the
 * master / reference frame before registration (c_image_stacking_pipeline.cc:1302-1306; defaults sigma 1, alpha 0.8).
 * CV_32F; 1 to 4 channels for sigma <= 2 (create_lpass_image's exact branch), single-channel for its pyramid
 * approximation (sigma > 2).  outmax <= outmin: no clamp.
 * ------------------------------------------------------------------------------------------- */
SSK_API int ssk_unsharp_mask(const ssk_mat *src, ssk_mat *dst, double sigma, double alpha, double outmin, double outmax);

/* ---------------------------------------------------------------------------------------------
 * Finishing step of the stacking pass: average_pyramid_inpaint(src, mask, dst, dstmask, max_levels)
 * (core/proc/inpaint/average_pyramid_inpaint.cc:97-127; call site c_image_stacking_pipeline.cc:763-767 with
 * max_levels = 100).  src / dst: CV_32F, 1 to 4 channels, same size; mask / dstmask: CV_8UC1 (dstmask may be null).
 * A null mask or a mask without holes returns copies of the inputs, like the reference.
 * ------------------------------------------------------------------------------------------- */
SSK_API int ssk_average_pyramid_inpaint(const ssk_mat *src, const ssk_mat *mask, ssk_mat *dst, ssk_mat *dstmask, int max_levels);
/* c_frame_accumulation::compute(avg, mask, dscale) followed by average_pyramid_inpaint(avg, mask, avg, mask, max_levels)
 * without leaving the device: what c_image_stacking_pipeline.cc:742-767 does with the accumulator at the end of a run. */
SSK_API int ssk_acc_compute_inpainted(ssk_acc *h, ssk_mat *avg, ssk_mat *mask, double dscale, int max_levels);

/* ---------------------------------------------------------------------------------------------
 * Stream-ordered call chains for device-resident data.  The reference's operators are blocking calls on cv::Mat
 * (lpg -> cv::GaussianBlur -> c_frame_accumulation::add in the focus-stack loop, c_jdr_pipeline.cc:1184-1236 per frame);
 * every libssk call therefore returns with its work finished.  With ssk_set_stream_ordered(1), ssk_lpg, ssk_gaussian_blur,
 * ssk_acc_add and ssk_jdr_derotate_and_add return as soon as the work is enqueued WHEN all their matrices are
 * SSK_MEM_DEVICE: the library orders these calls on the device (each waits for the previous one, whatever stream it runs
 * on), the host runs ahead, and nothing is lost to per-call launch and wait latency.  Calls with host matrices, and
 * ssk_acc_compute / _compute_inpainted / _get_counters / _clear and the other stateless operators, wait for the chain as
 * before.  Before reading a device result yourself, or handing it to another handle (ssk_stack_*, ssk_reg_*, ssk_ecch_*),
 * call ssk_device_synchronize().  The mode is process-wide, the chain is per host thread; returns the previous mode.
 * ------------------------------------------------------------------------------------------- */
SSK_API int ssk_set_stream_ordered(int enable);
SSK_API int ssk_device_synchronize(void);

/* ---------------------------------------------------------------------------------------------
 * Jovian derotation map: compute_ellipsoid_zrotation_remap (core/proc/feature2d/ellipsoid.cc:206-277), called by
 * c_jovian_derotation_remap::compute_derotation_for_angle (c_jovian_derotation_remap.cc:47-60).
 * R1 = pose of the ellipsoid as imaged, R2 = target pose (row-major 3x3, XYZscreen = R * XYZplanet); ebox_angle_deg and
 * crop_box {x, y, width, height} are ellipsoid_bbox(center, A, B, C, R2).angle and ellipse_crop_box(ebox, size), which
 * the caller computes (scalar geometry, ellipsoid.cc:16-84, 299-328).  Outputs: rmap CV_32FC2 (identity outside the
 * disk, (-1,-1) on the hidden side), wmap CV_32FC1 (limb weight wscale*sqrt(1-r^2), remapped by rmap), rmask CV_8UC1.
 * ------------------------------------------------------------------------------------------- */
/* Scalar geometry of the same classes, on the host: build_ellipsoid_rotation(pose = {longitude_rotation, tilt_to_earth,
 * position_angle}) = Rz * Rx * Ry (ellipsoid.h:47-71, pose.h:18-58), and ellipsoid_bbox(center, A, B, C, R) +
 * ellipse_crop_box(ebox, image_size) (ellipsoid.cc:16-84, 279-328): ebox = {center.x, center.y, width, height, angle_deg}
 * as the float members of the cv::RotatedRect, crop_box = {x, y, width, height}. */
SSK_API int ssk_build_ellipsoid_rotation(const double pose[3], double R[9]);
SSK_API int ssk_ellipsoid_bbox(int rows, int cols, const double center[2], const double axes[3], const double R[9],
                               float ebox[5], int crop_box[4]);
SSK_API int ssk_ellipsoid_zrotation_remap(int rows, int cols, const double center[2], const double axes[3],
                                          const double R1[9], const double R2[9], double ebox_angle_deg,
                                          const int crop_box[4], double wscale,
                                          ssk_mat *rmap, ssk_mat *wmap, ssk_mat *rmask);

/* One frame of c_jdr_pipeline::derotate_and_average_frames (core/pipeline/c_jdr_pipeline/c_jdr_pipeline.cc:1184-1236) and of
 * c_sdr_pipeline::derotate_and_average_frames (core/pipeline/c_sdr_pipeline/c_sdr_pipeline.cc:1192-1246: the same statements
 * over c_saturn_derotation_remap, which differs from the Jovian class by its rotation period only),
 * after preproc_align_and_remap: derotation map for R_current -> R_target (ssk_ellipsoid_zrotation_remap arguments),
 * weight = limb weight * wscale, 0 below 1e-5 [, * lpg(frame) remapped with BORDER_TRANSPARENT when
 * enable_weighted_average], 1 outside the disk for the master frame, 0 under the frame mask, GaussianBlur(1, REPLICATE);
 * frame remapped in place (INTER_LINEAR, BORDER_TRANSPARENT); acc.add(frame, weight).  frame: CV_32FC1, mask: CV_8UC1 / NULL. */
SSK_API int ssk_jdr_derotate_and_add(ssk_acc *acc, const ssk_mat *frame, const ssk_mat *mask, const double center[2],
                                     const double axes[3], const double R_current[9], const double R_target[9],
                                     double ebox_angle_deg, const int crop_box[4], double wscale, int is_master,
                                     int enable_weighted_average, double lpg_k, double lpg_p, int lpg_dscale,
                                     int lpg_uscale);

/* ---------------------------------------------------------------------------------------------
 * The fused per-frame loop of c_image_stacking_pipeline::process_input_sequence
 * (c_image_stacking_pipeline.cc:1358-1862): weights -> register -> warp(frame, mask, weights) -> accumulate,
 * for a batch of frames per call.  This is the data-parallel hot path; results are identical to calling
 * ssk_local_variance_map / ssk_reg_register_frame / ssk_reg_remap / ssk_acc_add frame by frame.
 * ------------------------------------------------------------------------------------------- */
enum { SSK_STACK_AVERAGE = 0, SSK_STACK_WEIGHTED_AVERAGE = 1, SSK_STACK_BAYER_AVERAGE = 2 };
/* frame_upscale_option / frame_upscale_stage, c_image_stacking_pipeline.h:33-49 (same values). */
enum { SSK_UPSCALE_NONE = 0, SSK_UPSCALE_PYRUP = 1, SSK_UPSCALE_X15 = 2, SSK_UPSCALE_X30 = 3 };
enum { SSK_UPSCALE_AFTER_ALIGN = 1, SSK_UPSCALE_BEFORE_ALIGN = 2 };
typedef struct ssk_stack_options {
  ssk_registration_options registration;
  int32_t accumulation_method;
  int32_t sm_dscale, sm_kradius, sm_uscale;  /* c_frame_accumulation_options::sharpness_measure, c_image_stacking_pipeline.h:94-99 */
  int32_t enable_registration;               /* c_image_stacking_options::enable_registration */
  int32_t bayer_colorid;                     /* for SSK_STACK_BAYER_AVERAGE */
  int32_t max_batch;                         /* frames in flight per call (device scratch is sized for it) */
  int32_t generating_master_frame;           /* the master-frame pass (create_reference_frame): frames are remapped with
                                                ECC_BORDER_REFLECT101 instead of registration.border_mode (c_image_stacking_pipeline.cc:1644-1651) */
  int32_t upscale_option;                    /* c_frame_upscale_options::upscale_option (SSK_UPSCALE_*), c_image_stacking_pipeline.h:57-86 */
  int32_t upscale_stage;                     /* SSK_UPSCALE_AFTER_ALIGN: the registration map is up-scaled and the frames are stacked at
                                                the up-scaled size (c_image_stacking_pipeline.cc:1633-1660; never during the master-frame
                                                pass, :2093-2101).  SSK_UPSCALE_BEFORE_ALIGN is not fused: up-scale the frames with
                                                ssk_upscale_image before the call, as INTEGRATION.md shows */
} ssk_stack_options;

typedef struct ssk_stack ssk_stack;
SSK_API void ssk_stack_options_default(ssk_stack_options *o);
SSK_API int ssk_stack_create(const ssk_stack_options *opts, ssk_stack **out);
SSK_API int ssk_stack_destroy(ssk_stack *h);
SSK_API int ssk_stack_set_reference(ssk_stack *h, const ssk_mat *image, const ssk_mat *mask, int bpp);
/* Process `n` frames (all of the same geometry/type).  frames[i] may be host or device memory.
 * transforms_out / status_out (arrays of n, may be NULL) receive the per-frame registration results. */
SSK_API int ssk_stack_add_frames(ssk_stack *h, const ssk_mat *frames, int n, int bpp,
                                 ssk_transform *transforms_out, ssk_ecc_status *status_out);
/* Enqueue only (no host sync, results stay on the device): for steady-state throughput measurement. */
SSK_API int ssk_stack_add_frames_async(ssk_stack *h, const ssk_mat *frames, int n, int bpp);
/* Streaming form: enqueue 1..max_batch frames and return at once; *ticket names the chunk.  Host frames must stay valid
 * and unchanged until ssk_stack_wait(ticket) returns.  Uploads of the next call overlap the processing of this one. */
SSK_API int ssk_stack_submit(ssk_stack *h, const ssk_mat *frames, int n, int bpp, int64_t *ticket);
/* Waits for chunk `ticket` and returns its per-frame registration results (*n_out frames).  Only the last 4 chunks
 * are kept. */
SSK_API int ssk_stack_wait(ssk_stack *h, int64_t ticket, ssk_transform *transforms_out, ssk_ecc_status *status_out,
                           int capacity, int *n_out);
SSK_API int ssk_stack_sync(ssk_stack *h);
/* Device-frame calls leave the border-ring part of their warp+accumulate stage running on a side stream so that it
 * overlaps the registration of the next call.  ssk_stack_flush makes the handle's stream wait for it (no host
 * synchronisation): call it before recording a timing event on ssk_stack_stream().  ssk_stack_sync, _compute,
 * _accumulated_frames and _accumulator include it. */
SSK_API int ssk_stack_flush(ssk_stack *h);
/* A new run over the same reference frame: empties the accumulator and the frame count, like the fresh c_frame_accumulation a
 * pipeline run starts with (create_frame_accumulation, c_image_stacking_pipeline.cc:450-466).  Stream-ordered, no host sync. */
SSK_API int ssk_stack_reset(ssk_stack *h);
/* c_frame_accumulation::compute() of the pipeline's accumulator. */
SSK_API int ssk_stack_compute(ssk_stack *h, ssk_mat *avg, ssk_mat *mask);
/* ssk_stack_compute + average_pyramid_inpaint (c_image_stacking_pipeline.cc:742-767) */
SSK_API int ssk_stack_compute_inpainted(ssk_stack *h, ssk_mat *avg, ssk_mat *mask, int max_levels);
SSK_API int ssk_stack_accumulated_frames(ssk_stack *h);
/* ssk_acc_reduce for the pipeline's accumulator, ordered after everything the handle has enqueued (the end of a sharded
 * run: c_image_stacking_pipeline.cc:731-769 then reads the root's accumulator). */
SSK_API int ssk_stack_reduce(ssk_stack *h, void *nccl_comm, int root);
SSK_API ssk_acc *ssk_stack_accumulator(ssk_stack *h);
SSK_API ssk_reg *ssk_stack_registration(ssk_stack *h);
/* CUDA stream the handle launches on (cudaStream_t as void*), for event timing by the caller. */
SSK_API void *ssk_stack_stream(ssk_stack *h);
/* Per-stage device time (ms) of the last ssk_stack_add_frames* call after ssk_stack_sync:
 * [0] prep (convert+pyrDown+smooth+pyramid) [1] weights [2] ECC [3] warp+accumulate.
 * The host-side enqueue of the same stages is bracketed by NVTX ranges ("ssk_stack: registration prep", "... sharpness
 * weights", "... register_frame (ECC)", "... warp + accumulate") for timeline tools. */
SSK_API int ssk_stack_stage_times(ssk_stack *h, float ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* SSK_H_ */
