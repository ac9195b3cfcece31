// Shared device/host helpers for the SerStacker hot-path kernels (sm_100a).
//
// The arithmetic here is the specification restated in oracle/cvmodel.py (checked against cv2 4.13):
// cv::remap's 1/32-px coordinate quantisation, its float bilinear / bicubic (A = -0.75) tables, its 15-bit
// fixed-point 8U tables, cv::borderInterpolate, and the reference's parametric maps
// (core/proc/image_registration/c_image_transform.cc) evaluated in the reference's operand order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/ssk.h"

namespace ssk {

// ------------------------------------------------------------------------------------------------
// errors / bookkeeping (host)
// ------------------------------------------------------------------------------------------------
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);
// stream-ordered call chains (ssk_runtime.cu)
int stream_ordered();
int set_stream_ordered(int enable);
int chain_wait(cudaStream_t s);
int chain_finish(cudaStream_t s, bool all_device);
int chain_drain();
int stream_after(cudaStream_t waiter, cudaStream_t producer);

#define SSK_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return ssk::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define SSK_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    ssk::count_launch();                                                      \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) return ssk::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define SSK_REQUIRE(cond, msg)                                                \
  do {                                                                        \
    if (!(cond)) { ssk::set_error(std::string(msg) + " [" #cond "]"); return SSK_ERR_INVALID; } \
  } while (0)

// ------------------------------------------------------------------------------------------------
// constants of cv::remap
// ------------------------------------------------------------------------------------------------
constexpr int kInterBits = 5;
constexpr int kInterTab = 32;
constexpr int kCoefBits = 15;
constexpr int kCoefScale = 1 << kCoefBits;

// tables built once per process on the host (ssk_tables.cu) and kept in device global memory
struct Tables {
  const float4 *cubic;      // [32] Keys coefficients at f/32 (float, A=-0.75)
  const short *cubic_itab;  // [32*32][16] fixed-point 2-D cubic weights, sum forced to 32768
  const float *lanczos;     // [32][8] interpolateLanczos4 coefficients at f/32 (float)
  const short *lanczos_itab;  // [32*32][64] fixed-point 2-D Lanczos weights, sum forced to 32768
};
int get_tables(Tables *t);   // lazily initialises for the current device

// ------------------------------------------------------------------------------------------------
// image views
// ------------------------------------------------------------------------------------------------
struct Img {           // read-only image on the device
  const void *data;
  int64_t step;        // bytes
  int rows, cols;
  int depth;           // SSK_8U / SSK_16U / SSK_32F
  int cn;
  float scale;         // multiplier applied on load (1/(1<<bpp) for integer frames, 1 for float)
};

template <int DEPTH> struct PixT;
template <> struct PixT<SSK_8U> { typedef uint8_t type; };
template <> struct PixT<SSK_16U> { typedef uint16_t type; };
template <> struct PixT<SSK_32F> { typedef float type; };

template <int DEPTH>
__device__ __forceinline__ float load_px(const Img &im, int y, int x, int c) {
  typedef typename PixT<DEPTH>::type T;
  const T *row = reinterpret_cast<const T *>(static_cast<const char *>(im.data) + (int64_t)y * im.step);
  if (DEPTH == SSK_32F) return __ldg(reinterpret_cast<const float *>(row) + x * im.cn + c);
  return __fmul_rn((float)__ldg(row + x * im.cn + c), im.scale);
}

// ------------------------------------------------------------------------------------------------
// cv::borderInterpolate
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int border_idx(int p, int n, int border) {
  if ((unsigned)p < (unsigned)n) return p;
  if (border == SSK_BORDER_REPLICATE || border == SSK_BORDER_TRANSPARENT) return p < 0 ? 0 : n - 1;
  if (border == SSK_BORDER_REFLECT || border == SSK_BORDER_REFLECT101) {
    const int delta = border == SSK_BORDER_REFLECT101 ? 1 : 0;
    if (n == 1) return 0;
    do {
      if (p < 0) p = -p - 1 + delta;
      else p = n - 1 - (p - n) - delta;
    } while ((unsigned)p >= (unsigned)n);
    return p;
  }
  if (border == SSK_BORDER_WRAP) {
    if (p < 0) p -= ((p - n + 1) / n) * n;
    if (p >= n) p %= n;
    return p;
  }
  return -1;  // CONSTANT
}

// cvRound(v * 32): integer part and 5-bit fraction
__device__ __forceinline__ void quant32(float v, int &ip, int &fp) {
  const int s = __float2int_rn(__fmul_rn(v, 32.0f));
  ip = s >> kInterBits;
  fp = s & (kInterTab - 1);
}

// ------------------------------------------------------------------------------------------------
// parametric maps: (x, y) -> (u, v), reference operand order, no FMA contraction
// ------------------------------------------------------------------------------------------------
enum { MAP_TRANSLATION = 0, MAP_EUCLIDEAN = 1, MAP_AFFINE = 3, MAP_HOMOGRAPHY = 4 };

struct MapCoef {
  int type;
  float c[9];
  // translation: c0=tx c1=ty
  // euclidean  : c0=scale c1=ca c2=sa c3=Tx c4=Ty c5=Cx c6=Cy
  // affine     : a00 a01 a02 a10 a11 a12
  // homography : a00..a22
};

__host__ __device__ inline MapCoef make_mapcoef(const ssk_transform &t) {
  MapCoef m;
  for (int i = 0; i < 9; ++i) m.c[i] = 0.f;
  switch (t.motion_type) {
    case SSK_MOTION_TRANSLATION:
      m.type = MAP_TRANSLATION; m.c[0] = t.params[0]; m.c[1] = t.params[1];
      break;
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN: {
      // parameters (tx, ty, angle[, scale]); fixed scale lives in aux[3] (c_image_transform.cc:384-423)
      const float angle = t.params[2];
      const float scale = t.motion_type == SSK_MOTION_SCALED_EUCLIDEAN ? t.params[3] : t.aux[3];
      m.type = MAP_EUCLIDEAN;
      m.c[0] = scale; m.c[1] = (float)cos((double)angle); m.c[2] = (float)sin((double)angle);
      m.c[3] = t.params[0]; m.c[4] = t.params[1]; m.c[5] = t.aux[0]; m.c[6] = t.aux[1];
      break;
    }
    case SSK_MOTION_AFFINE:
      m.type = MAP_AFFINE;
      for (int i = 0; i < 6; ++i) m.c[i] = t.params[i];
      break;
    default:
      m.type = MAP_HOMOGRAPHY;
      for (int i = 0; i < 8; ++i) m.c[i] = t.params[i];
      m.c[8] = t.aux[2];
      break;
  }
  return m;
}

__device__ __forceinline__ void map_xy(const MapCoef &m, float x, float y, float &u, float &v) {
  switch (m.type) {
    case MAP_TRANSLATION:   // c_image_transform.cc:158-167
      u = __fadd_rn(x, m.c[0]);
      v = __fadd_rn(y, m.c[1]);
      break;
    case MAP_EUCLIDEAN: {   // c_image_transform.cc:542-552
      const float xx = __fsub_rn(x, m.c[5]), yy = __fsub_rn(y, m.c[6]);
      u = __fadd_rn(__fmul_rn(m.c[0], __fsub_rn(__fmul_rn(m.c[1], xx), __fmul_rn(m.c[2], yy))), m.c[3]);
      v = __fadd_rn(__fmul_rn(m.c[0], __fadd_rn(__fmul_rn(m.c[2], xx), __fmul_rn(m.c[1], yy))), m.c[4]);
      break;
    }
    case MAP_AFFINE:        // c_image_transform.cc:935-943
      u = __fadd_rn(__fadd_rn(__fmul_rn(m.c[0], x), __fmul_rn(m.c[1], y)), m.c[2]);
      v = __fadd_rn(__fadd_rn(__fmul_rn(m.c[3], x), __fmul_rn(m.c[4], y)), m.c[5]);
      break;
    default: {              // c_image_transform.cc:1211-1220
      const float w = __fadd_rn(__fadd_rn(__fmul_rn(m.c[6], x), __fmul_rn(m.c[7], y)), m.c[8]);
      u = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.c[0], x), __fmul_rn(m.c[1], y)), m.c[2]), w);
      v = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(m.c[3], x), __fmul_rn(m.c[4], y)), m.c[5]), w);
      break;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cv::remap samplers on CV_32F-valued sources (float arithmetic, table weights, left-to-right sums)
// ------------------------------------------------------------------------------------------------
// Bilinear, one channel.  `border` as cv::BorderTypes; bval = constant border value.
template <int DEPTH>
__device__ __forceinline__ float sample_linear(const Img &im, int c, float u, float v, int border, float bval) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  const float tx = (float)fx * 0.03125f, ty = (float)fy * 0.03125f;
  const float wx[2] = {1.0f - tx, tx}, wy[2] = {1.0f - ty, ty};
  // cv::remapBilinear (scalar float path): ((S00*w00 + S01*w01) + S10*w10) + S11*w11, no contraction
  // (bit-exact against cv2 4.13, tests/test_cvmodel.py)
  const int y0 = border_idx(iy, im.rows, border), y1 = border_idx(iy + 1, im.rows, border);
  const int x0 = border_idx(ix, im.cols, border), x1 = border_idx(ix + 1, im.cols, border);
  const float s00 = (x0 >= 0 && y0 >= 0) ? load_px<DEPTH>(im, y0, x0, c) : bval;
  const float s01 = (x1 >= 0 && y0 >= 0) ? load_px<DEPTH>(im, y0, x1, c) : bval;
  const float s10 = (x0 >= 0 && y1 >= 0) ? load_px<DEPTH>(im, y1, x0, c) : bval;
  const float s11 = (x1 >= 0 && y1 >= 0) ? load_px<DEPTH>(im, y1, x1, c) : bval;
  float out = __fadd_rn(__fmul_rn(s00, __fmul_rn(wy[0], wx[0])), __fmul_rn(s01, __fmul_rn(wy[0], wx[1])));
  out = __fadd_rn(out, __fmul_rn(s10, __fmul_rn(wy[1], wx[0])));
  out = __fadd_rn(out, __fmul_rn(s11, __fmul_rn(wy[1], wx[1])));
  return out;
}

// Bicubic, one channel.
template <int DEPTH>
__device__ __forceinline__ float sample_cubic(const Img &im, int c, float u, float v, int border, float bval,
                                              const float4 *__restrict__ cubic) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  const float4 cx = __ldg(cubic + fx), cy = __ldg(cubic + fy);
  const float wx[4] = {cx.x, cx.y, cx.z, cx.w}, wy[4] = {cy.x, cy.y, cy.z, cy.w};
  float out = 0.f;
  const bool inside = ix >= 1 && iy >= 1 && ix + 2 < im.cols && iy + 2 < im.rows;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int yy = inside ? iy - 1 + ky : border_idx(iy - 1 + ky, im.rows, border);
    float row = 0.f;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int xx = inside ? ix - 1 + kx : border_idx(ix - 1 + kx, im.cols, border);
      const float s = (xx >= 0 && yy >= 0) ? load_px<DEPTH>(im, yy, xx, c) : bval;
      row = __fadd_rn(row, __fmul_rn(s, __fmul_rn(wy[ky], wx[kx])));
    }
    out = __fadd_rn(out, row);
  }
  return out;
}

// Lanczos4 (8 x 8 taps, first tap at ix - 3), one channel: the float weight of tap (ky, kx) is the product of the two 1-D
// table entries (initInterTab2D); remapLanczos4 adds the eight products of a row left to right and the rows top to bottom
// (bit-exact against cv2 4.13 where the 8 x 8 footprint is inside the image, oracle/cvmodel.py::remap_lanczos4_f32).
template <int DEPTH>
__device__ __forceinline__ float sample_lanczos4(const Img &im, int c, float u, float v, int border, float bval,
                                                 const float *__restrict__ lz) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  const float *wx = lz + fx * 8, *wy = lz + fy * 8;
  const bool inside = ix >= 3 && iy >= 3 && ix + 4 < im.cols && iy + 4 < im.rows;
  if (inside) {
    float out = 0.f;
#pragma unroll 1
    for (int ky = 0; ky < 8; ++ky) {
      const float wyk = __ldg(wy + ky);
      float row = 0.f;
#pragma unroll
      for (int kx = 0; kx < 8; ++kx) {
        const float term = __fmul_rn(load_px<DEPTH>(im, iy - 3 + ky, ix - 3 + kx, c), __fmul_rn(wyk, __ldg(wx + kx)));
        row = kx == 0 ? term : __fadd_rn(row, term);
      }
      out = __fadd_rn(out, row);
    }
    return out;
  }
  // footprint crossing the image edge: sum = cval + sum over the taps of (S - cval) w, tap by tap (remapLanczos4's border branch;
  // BORDER_TRANSPARENT positions its taps like BORDER_REFLECT_101, BORDER_CONSTANT skips the outside taps)
  const int tb = border == SSK_BORDER_TRANSPARENT ? SSK_BORDER_REFLECT101 : border;
  float sum = bval;
#pragma unroll 1
  for (int ky = 0; ky < 8; ++ky) {
    const int yy = border_idx(iy - 3 + ky, im.rows, tb);
    if (yy < 0) continue;
    const float wyk = __ldg(wy + ky);
#pragma unroll 1
    for (int kx = 0; kx < 8; ++kx) {
      const int xx = border_idx(ix - 3 + kx, im.cols, tb);
      if (xx >= 0) sum = __fadd_rn(sum, __fmul_rn(__fsub_rn(load_px<DEPTH>(im, yy, xx, c), bval), __fmul_rn(wyk, __ldg(wx + kx))));
    }
  }
  return sum;
}

// Nearest: cvRound of the coordinate.
template <int DEPTH>
__device__ __forceinline__ float sample_nearest(const Img &im, int c, float u, float v, int border, float bval) {
  const int ix = border_idx(__float2int_rn(u), im.cols, border);
  const int iy = border_idx(__float2int_rn(v), im.rows, border);
  return (ix >= 0 && iy >= 0) ? load_px<DEPTH>(im, iy, ix, c) : bval;
}

// ------------------------------------------------------------------------------------------------
// (cv::remap(Mat1b(size, 255), map, interp, BORDER_CONSTANT 0) >= thresh) for thresh in {250,254,255}.
// Bilinear: with an all-255 source the three thresholds coincide: valid iff every tap with a non-zero
// weight is in bounds (fixed-point weights are exact multiples of 32 for 1/32 fractions).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool valid255_linear(float u, float v, int cols, int rows) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  if (ix < 0 || iy < 0 || ix >= cols || iy >= rows) {
    // anchor outside: only valid if its weight is zero, which never happens (w00 = (32-fx)(32-fy) > 0)
    return false;
  }
  if (fx != 0 && ix + 1 >= cols) return false;
  if (fy != 0 && iy + 1 >= rows) return false;
  return true;
}

// Bicubic: value = saturate_u8((255 * sum_{taps in bounds} w + 2^14) >> 15) >= 255  <=>  sum >= 32704.
__device__ __forceinline__ bool valid255_cubic(float u, float v, int cols, int rows, const short *__restrict__ itab) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  if (ix >= 1 && iy >= 1 && ix + 2 < cols && iy + 2 < rows) return true;
  // sum of the in-bounds taps of the 4x4 fixed-point table: two 16-byte loads, byte-masked 16-bit dot products
  const uint4 *w = reinterpret_cast<const uint4 *>(itab + ((fy << kInterBits) + fx) * 16);
  const uint4 wa = __ldg(w), wb = __ldg(w + 1);
  unsigned cm = 0;
#pragma unroll
  for (int kx = 0; kx < 4; ++kx) cm |= ((unsigned)(ix - 1 + kx) < (unsigned)cols ? 1u : 0u) << (8 * kx);
  const int r0 = __dp2a_hi((int)wa.y, (int)cm, __dp2a_lo((int)wa.x, (int)cm, 0));
  const int r1 = __dp2a_hi((int)wa.w, (int)cm, __dp2a_lo((int)wa.z, (int)cm, 0));
  const int r2 = __dp2a_hi((int)wb.y, (int)cm, __dp2a_lo((int)wb.x, (int)cm, 0));
  const int r3 = __dp2a_hi((int)wb.w, (int)cm, __dp2a_lo((int)wb.z, (int)cm, 0));
  int S = 0;
  if ((unsigned)(iy - 1) < (unsigned)rows) S += r0;
  if ((unsigned)iy < (unsigned)rows) S += r1;
  if ((unsigned)(iy + 1) < (unsigned)rows) S += r2;
  if ((unsigned)(iy + 2) < (unsigned)rows) S += r3;
  int val = (255 * S + (1 << (kCoefBits - 1))) >> kCoefBits;
  return val >= 255;
}

// Lanczos4: as the bicubic rule with the 8 x 8 fixed-point table
static __device__ __noinline__ bool valid255_lanczos4(float u, float v, int cols, int rows, const short *__restrict__ itab) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  if (ix >= 3 && iy >= 3 && ix + 4 < cols && iy + 4 < rows) return true;
  const short *w = itab + ((fy << kInterBits) + fx) * 64;
  int S = 0;
  for (int ky = 0; ky < 8; ++ky) {
    if ((unsigned)(iy - 3 + ky) >= (unsigned)rows) continue;
    for (int kx = 0; kx < 8; ++kx)
      if ((unsigned)(ix - 3 + kx) < (unsigned)cols) S += __ldg(w + ky * 8 + kx);
  }
  return ((255 * S + (1 << (kCoefBits - 1))) >> kCoefBits) >= 255;
}

// itab: Tables::cubic_itab; Lanczos4 needs the whole Tables (valid255_t below)
__device__ __forceinline__ bool valid255(int interp, float u, float v, int cols, int rows, const short *itab) {
  if (interp == SSK_INTER_CUBIC) return valid255_cubic(u, v, cols, rows, itab);
  if (interp == SSK_INTER_NEAREST) {
    const int ix = __float2int_rn(u), iy = __float2int_rn(v);
    return (unsigned)ix < (unsigned)cols && (unsigned)iy < (unsigned)rows;
  }
  return valid255_linear(u, v, cols, rows);
}

__device__ __forceinline__ bool valid255_t(int interp, float u, float v, int cols, int rows, const Tables &tab) {
  if (interp == SSK_INTER_LANCZOS4) return valid255_lanczos4(u, v, cols, rows, tab.lanczos_itab);
  return valid255(interp, u, v, cols, rows, tab.cubic_itab);
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }
// cv::remap: `if (interpolation == INTER_AREA) interpolation = INTER_LINEAR`
inline int remap_interp(int v) { return v == SSK_INTER_AREA ? SSK_INTER_LINEAR : v; }

}  // namespace ssk
