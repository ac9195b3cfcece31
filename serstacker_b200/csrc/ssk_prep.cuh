// Internal launch interface of the per-frame preparation and weight-map kernels (ssk_prep.cu).
#pragma once
#include "ssk_common.cuh"

namespace ssk {

constexpr int kMaxTaps = 31;

// cv::pyrDown(src, dst, dstsize) ([1 4 6 4 1]/16 separable, BORDER_REFLECT_101, sample at 2x,2y).
// Source: any depth/cn (colour is converted to gray first: cv::cvtColor(COLOR_BGR2GRAY)); destination CV_32FC1 dense.
// Batched: src_ptrs[b] / dst_ptrs[b] are device arrays of per-frame pointers; if src_ptrs == null the single
// pointers src.data / dst are used.
struct PyrDownArgs {
  Img src;                       // geometry/type (data used when src_ptrs == null)
  const void *const *src_ptrs;   // device array [batch] or null
  float *dst; float *const *dst_ptrs;
  int dst_rows, dst_cols;
  int batch;
  float post_scale;              // multiplies the result (lpg: 1/(1+dscale)); 1 = none
  int border;                    // 0 (default): BORDER_REFLECT101 (cv::pyrDown's default); SSK_BORDER_REPLICATE: ecc_downscale
};
int launch_pyrdown(const PyrDownArgs &a, cudaStream_t s);

#ifdef __CUDACC__
// cv::cvtColor(COLOR_BGR2GRAY) on float data: 0.114 B + 0.587 G + 0.299 R (the gray load of the prep kernels)
__device__ __forceinline__ float bgr2gray(float b, float g, float r) { return fmaf(r, 0.299f, fmaf(g, 0.587f, b * 0.114f)); }
// cv::pyrDown's 1-4-6-4-1 forms as cv2 4.13 evaluates them (see k_pyrdown, ssk_prep.cu): `simd` selects the operand order of
// the vectorised loop, otherwise the scalar tail's.  pd_vform includes the final 1/256.
__device__ __forceinline__ float pd_hform(float p0, float p1, float p2, float p3, float p4, bool simd) {
  const float a1 = __fmul_rn(__fadd_rn(p1, p3), 4.f), c6 = __fmul_rn(p2, 6.f);
  return simd ? __fadd_rn(c6, __fadd_rn(a1, __fadd_rn(p0, p4))) : __fadd_rn(__fadd_rn(__fadd_rn(c6, a1), p0), p4);
}
__device__ __forceinline__ float pd_vform(float r0, float r1, float r2, float r3, float r4, bool simd) {
  const float a13 = __fadd_rn(r1, r3);
  float v;
  if (simd) v = __fadd_rn(__fmul_rn(__fadd_rn(a13, r2), 4.f), __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2)));
  else v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2, 6.f), __fmul_rn(a13, 4.f)), r0), r4);
  return __fmul_rn(v, 1.0f / 256.0f);
}
#endif

// cv::resize(INTER_AREA) for down-scaling (scale >= 1 in both axes), single-channel fp32 result (colour / integer frames go
// through the same gray load as pyrDown): scaleImage's branch for ecc.scale != 0.5 (c_frame_registration.cc:242-247).
// inv_scale_x/y are cv::resize's inv_scale_* (fx, fy when dsize was derived from them, else dsize / ssize).
struct ResizeAreaArgs {
  Img src; const void *const *src_ptrs;
  float *dst; float *const *dst_ptrs;
  int dst_rows, dst_cols;
  double inv_scale_x, inv_scale_y;
  int batch;
};
int launch_resize_area(const ResizeAreaArgs &a, cudaStream_t s);

// cv::pyrUp(src, dst, dstsize) on dense CV_32FC1 images (dst_cols in {2*cols - 1, 2*cols, 2*cols + 1}, same for rows),
// optionally fused with ecc_normalize's subtraction: dst = minuend - pyrUp(src), zeroed where mask == 0.
struct PyrUpArgs {
  const float *src; const float *const *src_ptrs;
  float *dst; float *const *dst_ptrs;
  int rows, cols, dst_rows, dst_cols, batch;
  const float *const *minuend_ptrs; const float *minuend;   // null: plain pyrUp
  const uint8_t *mask;                                      // dense dst-size mask shared by the batch, or null
};
int launch_pyrup(const PyrUpArgs &a, cudaStream_t s);

// cv::sepFilter2D(src, dst, CV_32F, kx, ky, anchor centre, BORDER_REPLICATE) on dense CV_32FC1 images.
struct SepFilterArgs {
  const float *src; const float *const *src_ptrs;
  float *dst; float *const *dst_ptrs;
  int rows, cols, batch;
  int kxn, kyn;                  // odd tap counts
  float kx[kMaxTaps], ky[kMaxTaps];
  int border;                    // 0 / SSK_BORDER_REPLICATE (default), SSK_BORDER_REFLECT, SSK_BORDER_REFLECT101
  int cn;                        // interleaved channels filtered independently (0 or 1: single channel)
};
int launch_sepfilter(const SepFilterArgs &a, cudaStream_t s);

// gray / depth conversion only (ecc.scale == 1): dst = CV_32FC1 dense
int launch_to_gray(const Img &src, const void *const *src_ptrs, float *dst, float *const *dst_ptrs, int batch,
                   cudaStream_t s);

// W1: compute_local_variance_map (c_local_variance_sharpness_measure.cc:193-247) on a CV_32FC1 image M
// (already pyrDown'ed `dscale` times):  G = morph-gradient(k x k, REPLICATE); map = G^3 s^3 + 0.05 Q;
// out = resize(map, full size, INTER_LINEAR).
struct W1Args {
  const float *M; const float *const *M_ptrs;     // [batch] small images, rows x cols
  int rows, cols;
  int kradius;
  double depth_scale;                             // 20 * 1 / maxval(depth)
  float *gmap; float *const *gmap_ptrs;           // scratch rows x cols per frame
  double *partials;                               // scratch [batch][2][nblocks]
  double *stats;                                  // [batch][4]: sumG, sumG4, Q, add
  float *out; float *const *out_ptrs;             // full_rows x full_cols per frame (dense)
  int full_rows, full_cols;
  int batch;
  int2 *axis_tab;                                 // scratch [full_cols + full_rows]: per output column / row source index + fraction
  // uscale > 0 (c_local_variance_sharpness_measure.cc:231-234): the map is reduced to dscaleSize(size, uscale) by
  // cv::resize(INTER_AREA) before the 0.05 Q offset and the up-sampling; gmap2 is scratch of rows x cols floats per frame
  int uscale; float *gmap2; float *const *gmap2_ptrs;
  int *axis_tab_built;                            // host flag (optional): the tables in axis_tab are already those of this geometry
};
int w1_num_blocks(int rows, int cols);
void w1_uscale_size(int rows, int cols, int uscale, int *urows, int *ucols);   // dscaleSize()
int launch_w1(const W1Args &a, cudaStream_t s);

// W2 (lpg.cc): 5x5 Laplacian/gradient energy; in-place scale + integer power
int launch_lpg5x5(const float *src, int rows, int cols, float *dst, float alpha, float beta, float eps, cudaStream_t s);
// channel average + 5x5 operator + integer power in one pass (the dscale = 0 form of lpg: no scaling before the operator)
int launch_lpg_fused(const Img &im, float *dst, float alpha, float beta, float eps, int ipow, cudaStream_t s);
int launch_scale_ipow(float *buf, int64_t n, float scale, bool apply_scale, int ipow, cudaStream_t s);
// the small levels of lpg's pyramid (down, scale, power, up) in one 8-CTA cluster launch; P in place, Q scratch
int launch_lpg_tail(float *P, float *Q, int rows, int cols, int ndown, float scale, int ipow, cudaStream_t s);

// reference-mask helpers: 8U pyrDown + threshold, INTER_NEAREST resize, gradient masking + non-zero count
int launch_pyrdown_mask_u8(const uint8_t *src, int64_t sstep, int rows, int cols, uint8_t *dst, int drows, int dcols, int thresh,
                           cudaStream_t s);
int launch_resize_nearest_u8(const uint8_t *src, int rows, int cols, uint8_t *dst, int drows, int dcols, cudaStream_t s);
int launch_apply_refmask(const uint8_t *mask, int n, float *gx, float *gy, int *d_count, cudaStream_t s);

// erode 5x5 / 8U helpers for user masks
int launch_erode5_u8(const uint8_t *src, int64_t sstep, uint8_t *dst, int64_t dstep, int rows, int cols,
                     int border_replicate, cudaStream_t s);

// unsharp_mask's combine step (unsharp_mask.cc:106): dst = src * alpha + lpass * beta as cv::addWeighted computes it for
// CV_32F (fp64 fma(src, alpha, lpass * beta) rounded to fp32), then the optional cv::min / cv::max clamp.  n = all samples.
int launch_add_weighted(const float *src, double alpha, const float *lpass, double beta, float *dst, int64_t n, int clamp,
                        float outmin, float outmax, cudaStream_t s);

// debayer_nn2 (core/io/debayer.cc:827-1195; ssk_debayer.cu): raw Bayer CV_8U / CV_16U / CV_32F single channel -> BGR of the
// same depth; both images on the device.
int launch_debayer_nn2(const void *src, int64_t sstep, int depth, int rows, int cols, int colorid, void *dst, int64_t dstep,
                       cudaStream_t s);
// debayer_nn2 -> BGR2GRAY -> cv::pyrDown fused (the ECC image of a raw Bayer frame at ecc.scale 0.5), bit-identical to the chain
int launch_bayer_gray_pyrdown(const void *const *src_ptrs, bool frames_aligned16, int64_t sstep, int depth, int rows, int cols, int colorid, float scale,
                              float *const *dst_ptrs, int batch, cudaStream_t s);

// average_pyramid_inpaint (core/proc/inpaint/average_pyramid_inpaint.cc:97-127; ssk_inpaint.cu).  src: CV_32F with `cn`
// interleaved channels, mask: CV_8UC1 (both on the device, any step); dst / dstmask dense; `work` holds
// inpaint_work_bytes() bytes.  *was_full = 1 when the mask had no holes (outputs are copies of the inputs).
size_t inpaint_work_bytes(int rows, int cols, int cn, int max_levels);
// linear_interpolation_inpaint (core/proc/inpaint/linear_interpolation_inpaint.cc; ssk_lininpaint.cu), in place on dense device buffers
size_t lin_inpaint_work_bytes(int rows, int cols);
int launch_linear_interpolation_inpaint(float *img, int64_t istep, uint8_t *mask, int64_t mstep, int rows, int cols, int cn, void *work,
                                        cudaStream_t s);
int launch_average_pyramid_inpaint(const float *src, int64_t sstep, const uint8_t *mask, int64_t mstep, int rows, int cols, int cn,
                                   int max_levels, void *work, float *dst, uint8_t *dstmask, int *was_full, cudaStream_t s);

}  // namespace ssk
