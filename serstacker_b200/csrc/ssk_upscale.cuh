// frame_upscale_option of c_image_stacking_pipeline (c_image_stacking_pipeline.h:57-86, c_image_stacking_pipeline.cc:1869-2002):
// x2.0 = cv::pyrUp, x1.5 = cv::resize(INTER_LINEAR) to (3 w / 2, 3 h / 2), x3.0 = cv::resize(INTER_LINEAR_EXACT) to (3 w, 3 h)
// (for CV_32F data OpenCV evaluates INTER_LINEAR_EXACT as INTER_LINEAR).  Device evaluators of one destination sample from a
// source functor S(x, y) -> float, so that an up-scaled registration map is evaluated on the fly and never materialised.
#pragma once
#include "ssk_common.cuh"

namespace ssk {

inline void upscale_size(int option, int cols, int rows, int *ucols, int *urows) {
  switch (option) {
    case SSK_UPSCALE_PYRUP: *ucols = cols * 2; *urows = rows * 2; break;
    case SSK_UPSCALE_X15: *ucols = cols * 3 / 2; *urows = rows * 3 / 2; break;
    case SSK_UPSCALE_X30: *ucols = cols * 3; *urows = rows * 3; break;
    default: *ucols = cols; *urows = rows; break;
  }
}

// one axis of cv::resize(INTER_LINEAR) on float data: fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), clamped at both
// ends with a zero fraction (resize.cpp: "if (sx < 0) fx = 0, sx = 0; if (sx >= ssize - 1) fx = 0, sx = ssize - 1")
struct LinAxis { int s0, s1; float a0, a1; };
__device__ __forceinline__ LinAxis lin_axis(int d, int ssize, double scale) {
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(fx);
  fx -= (float)s;
  if (s < 0) { fx = 0.f; s = 0; }
  if (s >= ssize - 1) { fx = 0.f; s = ssize - 1; }
  LinAxis a;
  a.s0 = s; a.s1 = min(s + 1, ssize - 1); a.a0 = 1.f - fx; a.a1 = fx;
  return a;
}

// cv::resize(INTER_LINEAR) sample: horizontal pass of the two source rows (S0 a0 + S1 a1), then the vertical mix
template <class SRC>
__device__ __forceinline__ float up_linear(SRC S, const LinAxis &ax, const LinAxis &ay) {
  const float r0 = __fadd_rn(__fmul_rn(S(ax.s0, ay.s0), ax.a0), __fmul_rn(S(ax.s1, ay.s0), ax.a1));
  const float r1 = __fadd_rn(__fmul_rn(S(ax.s0, ay.s1), ax.a0), __fmul_rn(S(ax.s1, ay.s1), ax.a1));
  return __fadd_rn(__fmul_rn(r0, ay.a0), __fmul_rn(r1, ay.a1));
}

// cv::pyrUp sample (dst = 2 w x 2 h), the arithmetic of ssk_prep.cu::k_pyrup (bit-exact against cv2, tests/test_cvmodel.py)
template <class SRC>
__device__ __forceinline__ float up_pyr_row(SRC S, int w, int ox, int y) {
  const int x = min(ox >> 1, w - 1);
  const bool odd = (ox & 1) || (ox >> 1) >= w;
  if (odd) return x == w - 1 ? __fmul_rn(S(x, y), 8.f) : __fmul_rn(__fadd_rn(S(x, y), S(x + 1, y)), 4.f);
  if (x == 0) return __fadd_rn(__fmul_rn(S(0, y), 6.f), __fmul_rn(S(1, y), 2.f));
  if (x == w - 1) return __fadd_rn(S(x - 1, y), __fmul_rn(S(x, y), 7.f));
  return __fadd_rn(__fadd_rn(S(x - 1, y), __fmul_rn(S(x, y), 6.f)), S(x + 1, y));
}
template <class SRC>
__device__ __forceinline__ float up_pyr(SRC S, int w, int h, int ox, int oy) {
  const int y = min(oy >> 1, h - 1);
  const bool odd = (oy & 1) || (oy >> 1) >= h;
  const int yd = min(y + 1, h - 1), yu = y == 0 ? 1 : y - 1;
  const float r1 = up_pyr_row(S, w, ox, y), r2 = up_pyr_row(S, w, ox, yd);
  float v;
  if (odd) v = __fmul_rn(__fadd_rn(r1, r2), 4.f);
  else v = __fadd_rn(__fadd_rn(__fmul_rn(r1, 6.f), up_pyr_row(S, w, ox, yu)), r2);
  return __fmul_rn(v, 1.0f / 64.0f);
}

// geometry of one up-scaling, passed to kernels by value
struct UpscaleGeom {
  int option;
  int sw, sh, dw, dh;
  double scale_x, scale_y;      // cv::resize: 1 / (dsize / ssize)
};
inline UpscaleGeom make_upscale_geom(int option, int cols, int rows) {
  UpscaleGeom g;
  g.option = option; g.sw = cols; g.sh = rows;
  upscale_size(option, cols, rows, &g.dw, &g.dh);
  g.scale_x = 1.0 / ((double)g.dw / cols); g.scale_y = 1.0 / ((double)g.dh / rows);
  return g;
}

template <class SRC>
__device__ __forceinline__ float upscale_sample(const UpscaleGeom &g, SRC S, int ox, int oy) {
  if (g.option == SSK_UPSCALE_PYRUP) return up_pyr(S, g.sw, g.sh, ox, oy);
  if (g.option == SSK_UPSCALE_NONE) return S(ox, oy);
  return up_linear(S, lin_axis(ox, g.sw, g.scale_x), lin_axis(oy, g.sh, g.scale_y));
}

// Up-scaling of a binary CV_8UC1 mask followed by `>= 255` (upscale_image, c_image_stacking_pipeline.cc:1996-1998): the 8-bit
// interpolators give 255 exactly where every source tap that carries weight is 255 (the smallest weight of a zero tap - 1/6 for
// x1.5, 1/3 for x3.0, 1/64 for pyrUp - already pulls the value below 254.5)
template <class MSK>
__device__ __forceinline__ bool upscale_mask(const UpscaleGeom &g, MSK M, int ox, int oy) {
  if (g.option == SSK_UPSCALE_NONE) return M(ox, oy);
  if (g.option == SSK_UPSCALE_PYRUP) {
    const int x = min(ox >> 1, g.sw - 1), y = min(oy >> 1, g.sh - 1);
    const bool oddx = (ox & 1) || (ox >> 1) >= g.sw, oddy = (oy & 1) || (oy >> 1) >= g.sh;
    const int x0 = oddx ? x : max(x - 1, 0), x1 = min(x + 1, g.sw - 1);
    const int y0 = oddy ? y : max(y - 1, 0), y1 = min(y + 1, g.sh - 1);
    bool ok = true;
    for (int yy = y0; yy <= y1; ++yy)
      for (int xx = x0; xx <= x1; ++xx) ok = ok && M(xx, yy);
    return ok;
  }
  const LinAxis ax = lin_axis(ox, g.sw, g.scale_x), ay = lin_axis(oy, g.sh, g.scale_y);
  if (g.option == SSK_UPSCALE_X15) {
    // cv::resize(INTER_LINEAR) on CV_8U: 11-bit fixed-point coefficients (cvRound(f * 2048)), horizontal pass in int,
    // vertical pass ((b0 (S0 >> 4)) >> 16) + ((b1 (S1 >> 4)) >> 16) + 2) >> 2 (resize.cpp, HResizeLinear / VResizeLinear<uchar>):
    // a zero tap whose coefficient rounds to 0 or 1 / 2048 still leaves 255 (sizes that are not a multiple of 2)
    const int ia0 = __float2int_rn(ax.a0 * 2048.f), ia1 = __float2int_rn(ax.a1 * 2048.f);
    const int ib0 = __float2int_rn(ay.a0 * 2048.f), ib1 = __float2int_rn(ay.a1 * 2048.f);
    const int h0 = (M(ax.s0, ay.s0) ? 255 : 0) * ia0 + (M(ax.s1, ay.s0) ? 255 : 0) * ia1;
    const int h1 = (M(ax.s0, ay.s1) ? 255 : 0) * ia0 + (M(ax.s1, ay.s1) ? 255 : 0) * ia1;
    return ((((ib0 * (h0 >> 4)) >> 16) + ((ib1 * (h1 >> 4)) >> 16) + 2) >> 2) >= 255;
  }
  bool ok = M(ax.s0, ay.s0);
  if (ax.a1 != 0.f) ok = ok && M(ax.s1, ay.s0);
  if (ay.a1 != 0.f) ok = ok && M(ax.s0, ay.s1);
  if (ax.a1 != 0.f && ay.a1 != 0.f) ok = ok && M(ax.s1, ay.s1);
  return ok;
}

// launchers (ssk_upscale.cu): dense device images
int launch_upscale_f32(int option, const float *src, int64_t sstep, int rows, int cols, int cn, float *dst, int64_t dstep, float post_scale, cudaStream_t s);
int launch_upscale_mask(int option, const uint8_t *src, int64_t sstep, int rows, int cols, uint8_t *dst, int64_t dstep, cudaStream_t s);

}  // namespace ssk
