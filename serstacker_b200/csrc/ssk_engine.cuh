// Host-side engine classes behind the C ABI (ssk_api.cu): device-resident equivalents of the reference's
// c_ecch, c_frame_registration, c_frame_accumulation and the per-frame loop of c_image_stacking_pipeline.
#pragma once
#include <vector>
#include "ssk_common.cuh"
#include "ssk_ecc.cuh"
#include "ssk_prep.cuh"
#include "ssk_warp.cuh"

namespace ssk {

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  int ensure(size_t n);          // (re)allocates when n > bytes
  void release();
  template <class T> T *as() const { return static_cast<T *>(p); }
};

struct PinnedBuf {
  void *p = nullptr;
  size_t bytes = 0;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  int ensure(size_t n);
  template <class T> T *as() const { return static_cast<T *>(p); }
};

inline int type_depth(int type) { return type & 7; }
inline int type_cn(int type) { return (type >> 3) + 1; }
inline int depth_bytes(int depth) { return depth == SSK_8U ? 1 : depth == SSK_16U ? 2 : depth == SSK_32F ? 4 : 0; }
inline float bpp_scale(int depth, int bpp) {
  // c_image_stacking_pipeline_base.cc:271-276: convertTo(CV_32F, 1/(1<<bpp)) for integer frames only
  return depth == SSK_32F ? 1.0f : (float)(1.0 / (double)(1 << bpp));
}

// A frame made available on the device (copied from the host if needed).
struct DeviceImage {
  const void *data = nullptr;
  int64_t step = 0;
  int rows = 0, cols = 0, type = 0;
};

// ------------------------------------------------------------------------------------------------
// c_ecch: reference pyramid + batched alignment of current images
// ------------------------------------------------------------------------------------------------
// Validates a CV_8UC1 mask of the given size and makes it available on the device (host data is staged, dense).
int mask_to_device(const ssk_mat *mask, int rows, int cols, DevBuf &staging, cudaStream_t s, const uint8_t **d_mask, int64_t *step);

class Ecch {
 public:
  ssk_ecch_options opts;
  cudaStream_t stream = nullptr;
  bool cluster_fixed = false;   // SSK_ECC_CLUSTER given
  int cluster_size = 8;   // CTAs per frame (SSK_ECC_CLUSTER overrides); 8 x 256 threads, >= 2 CTAs per SM: a 64-frame batch is resident at once

  int nlevels = 0;
  int lw[kMaxLevels], lh[kMaxLevels];
  int64_t loff[kMaxLevels];        // float offset of each level in a pyramid buffer
  int64_t pyr_floats = 0;
  bool have_reference = false;
  bool first_align_pending = false;  // parameter-dependent steepest-descent images not captured yet
  int hp_main_type = -1;             // motion type the cached main-transform Hp was built for

  // registration flow (set by Reg; zero for bare c_ecch use)
  int motion_type = SSK_MOTION_TRANSLATION;
  int translation_first = 0, check_rho = 0;
  double min_rho = 0, final_scale = 1;

  ~Ecch();
  int init(const ssk_ecch_options &o, cudaStream_t s);
  // reference: dense CV_32FC1 device image of level-0 size (already scaled to the ECC resolution)
  // d_mask: dense CV_8UC1 reference mask of the same size on the device, or null
  int set_reference(const float *d_img, int rows, int cols, const uint8_t *d_mask = nullptr);
  // scratch for `batch` frames in flight
  int reserve(int batch);
  // level-0 source images (dense CV_32FC1, device) of the frames of a batch -> smoothed pyramids
  int prepare_current(const float *const *d_src_ptrs /*device array*/, int batch);
  // current mask of the frame in slot 0 (dense CV_8UC1 of level-0 size on the device): pyramid by cv::resize(INTER_NEAREST)
  // (c_ecch::set_current_image, ecc2.cc:1060-1120), eroded 5x5 per level for the forward-additive solver
  // (ecc2.cc:1214-1234).  Applies to the next align() only, which must be a single-frame one.
  int prepare_current_mask(const uint8_t *d_mask);
  // run the alignment for `batch` prepared frames, every frame starting from t0
  int align(int batch, const ssk_transform &t0);
  EccFrame *device_frames() { return d_frames.as<EccFrame>(); }
  EccFrame *host_frames() { return h_frames.as<EccFrame>(); }
  int download_frames(int batch);   // async copy of EccFrame records to the pinned mirror
  const float *reference_level(int l) const { return ref_pyr.as<float>() + loff[l]; }
  const float *current_level(int slot, int l) const { return cur_pyr.as<float>() + (int64_t)slot * pyr_floats + loff[l]; }
  float *level0_scratch(int slot) { return src0.as<float>() + (int64_t)slot * lw[0] * lh[0]; }
  float *const *level0_scratch_ptrs() { return d_src0_ptrs.as<float *>(); }
  int capacity = 0;
  // debug trace of frame 0 of every align (kTraceRec floats per solver trial)
  int enable_trace(int max_records);
  int fetch_trace(float *out, int max_records, int *n);

 private:
  int build_config();
  int hp_mode_for_next_align() const;
  DevBuf ref_pyr, ref_gx, ref_gy, cur_pyr, src0, tmp;
  DevBuf cur_mask, cur_mask_tmp;
  bool cur_mask_pending = false;
  DevBuf ref_mask, ref_mask_tmp, d_count;          // reference-mask pyramid (bytes, level l at loff[l]), erode scratch, counters
  bool have_ref_mask = false;
  double rma[kMaxLevels];                          // reference mask area per level
  DevBuf d_hp_trans, d_hp_main, d_frames, d_trace;
  int trace_capacity = 0;
  DevBuf d_lvl_ptrs, d_src0_ptrs, d_tmp_ptrs;    // per-level pointer tables for the batched kernels
  PinnedBuf h_frames;
  EccConfig cfg;
  float gauss_ref[kMaxTaps], gauss_cur[kMaxTaps];
  int gauss_ref_n = 0, gauss_cur_n = 0;
};

// ------------------------------------------------------------------------------------------------
// c_weigthed_average / c_bayer_average
// ------------------------------------------------------------------------------------------------
class Acc {
 public:
  int kind = SSK_ACC_WEIGHTED_AVERAGE;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int rows = 0, cols = 0, cn = 0;
  int frames = 0;
  int colorid = SSK_COLORID_BAYER_RGGB;
  bool have_map = false;
  ssk_transform map_t;
  DevBuf rmap;       // explicit CV_32FC2 map (dense) when set_remap was given one
  bool rmap_explicit = false;
  int rmap_rows = 0, rmap_cols = 0;
  DevBuf acc, wacc;  // mean (rows*cols*cn) + weights (rows*cols)   |  bayer: acc (x3) + counters (x3)
  DevBuf staging, wstaging, out_staging;
  ~Acc();
  int ensure(int rows, int cols, int cn);   // allocate + zero on first use
  int clear();
};

// ------------------------------------------------------------------------------------------------
// c_frame_registration (ECC branch)
// ------------------------------------------------------------------------------------------------
class EccFlow;                         // ssk_eccflow.cuh
struct EccFlowHolder {                 // owning pointer with the deleter out of line (EccFlow is incomplete here)
  EccFlow *p = nullptr;
  EccFlowHolder() = default;
  EccFlowHolder(const EccFlowHolder &) = delete;
  EccFlowHolder &operator=(const EccFlowHolder &) = delete;
  ~EccFlowHolder();
  EccFlow *operator->() const { return p; }
};

class Reg {
 public:
  ssk_registration_options opts;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  Ecch ecch;
  EccFlowHolder flowh;                 // c_eccflow stage (enable_eccflow_registration), null when disabled
  bool flow_enabled() const { return opts.enable_eccflow_registration != 0; }
  int ref_rows = 0, ref_cols = 0;      // _reference_frame_size
  int ecc_rows = 0, ecc_cols = 0;      // size of the ECC image (after scaleImage)
  bool have_current = false;
  ssk_transform current;               // result of the last register_frame
  ssk_transform default_transform;     // _image_transform_defaut_parameters
  DevBuf staging, mask_tmp, out_staging, flow_img, flow_mask;
  DevBuf d_one_ptr;                    // 1-entry pointer tables for the single-frame path
  DevBuf norm_buf, norm_ptrs;          // ecc_normalize: per-frame pyrDown chain + per-level pointer tables
  int norm_capacity = 0;
  // scaleImage took the cv::pyrDown branch (ecc.scale == 0.5): the ECC image is also the first level of W1's pdownscale chain
  bool scaled_by_pyrdown() const { return opts.ecc.scale > 0 && opts.ecc.scale > 0.49 && opts.ecc.scale < 0.51; }
  bool normalize_enabled() const { return opts.ecc.normalization_scale > 0 && opts.ecc.normalization_noise > 0; }
  // ecc_normalize (ecc2.cc:385-397) in place on `batch` dense ecc_rows x ecc_cols images; d_mask: dense mask or null
  int normalize(float *const *d_img_ptrs, int batch, const uint8_t *d_mask);
  ~Reg();
  int init(const ssk_registration_options &o, cudaStream_t s, bool own);
  // d_mask / mask_step: full-resolution CV_8UC1 reference mask on the device, or null
  int setup_reference(const Img &frame, const uint8_t *d_mask = nullptr, int64_t mask_step = 0);
  // frames (device, common geometry) -> ECC images -> pyramids.  d_frame_ptrs: device array of frame pointers.
  // d_mask / mask_step: full-resolution CV_8UC1 mask of the (single) current frame on the device, or null
  int prepare(const Img &geom, const void *const *d_frame_ptrs, int batch, const uint8_t *d_mask = nullptr, int64_t mask_step = 0);
  // bayer_average loop: reserve the batch's slots so that the caller can write the scaled ECC images of raw Bayer frames
  // straight into ecch.level0_scratch_ptrs() (launch_bayer_gray_pyrdown), then prepare(geom, nullptr, batch) with
  // ecc_images_ready set skips scaleImage
  int reserve_batch(int batch);
  bool ecc_images_ready = false;
  int register_batch(int batch);       // launches the ECC kernel; results in ecch.device_frames()
};

int make_transform(ssk_transform *t, int motion_type);
int host_scale_transform(ssk_transform *t, double f);

}  // namespace ssk

// ------------------------------------------------------------------------------------------------
// opaque handle types of include/ssk.h
// ------------------------------------------------------------------------------------------------
struct ssk_ecch {
  ssk::Ecch e;
  cudaStream_t stream = nullptr;
  ssk::DevBuf staging, d_ptr, st_mask, st_mask2;
  ~ssk_ecch() { if (stream) cudaStreamDestroy(stream); }
};

struct ssk_reg {
  ssk::Reg r;
  ssk::DevBuf staging, st_map, st_mask, st_out, st_tmp, d_ptr, st_flowmap;
};

struct ssk_acc {
  ssk::Acc a;
};
