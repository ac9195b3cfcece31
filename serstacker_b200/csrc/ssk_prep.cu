// K1 (per-frame ECC preparation) and K6 (W1 sharpness weight map) kernels.
//
// Reference semantics reproduced here:
//   scaleImage (cv::pyrDown of the ECC image)          core/proc/image_registration/c_frame_registration.cc:230-250
//   c_ecch::set_current_image / set_reference_image    core/proc/image_registration/ecc2.cc:984-1120
//        (Gaussian sepFilter2D BORDER_REPLICATE, pyramid by cv::pyrDown(img, nextSize))
//   extract_channel(gray) = cv::cvtColor(BGR2GRAY)     core/proc/extract_channel.cc:607-697
//   compute_local_variance_map                         core/proc/sharpness_measure/c_local_variance_sharpness_measure.cc:193-247
// All kernels are batched over the frames of a batch through blockIdx.z.
#include "ssk_prep.cuh"
#include <cooperative_groups.h>
#include <cfloat>
#include <cmath>

namespace ssk {

namespace {

template <int DEPTH>
__device__ __forceinline__ float load_gray(const Img &im, int y, int x) {
  if (im.cn == 1) return load_px<DEPTH>(im, y, x, 0);
  // cv::cvtColor(COLOR_BGR2GRAY) on float data: 0.114 B + 0.587 G + 0.299 R
  const float b = load_px<DEPTH>(im, y, x, 0), g = load_px<DEPTH>(im, y, x, 1), r = load_px<DEPTH>(im, y, x, 2);
  return bgr2gray(b, g, r);
}

// Visits the (ih x iw) elements of a CTA tile with one warp per row and lanes along x (coalesced, no div/mod).
template <class F>
__device__ __forceinline__ void for_tile(int ih, int iw, F f) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < ih; r += nw)
    for (int c = lane; c < iw; c += 32) f(r, c);
}

constexpr int PD_OW = 64, PD_OH = 16, PD_RPT = 4;   // CTA tile (outputs) and output rows per thread

// cv::pyrDown, bit-exact against cv2 4.13 (oracle/cvmodel.py::pyrdown_f32).
//   horizontal: columns covered by the 4-lane PyrDownVecH loop use s0*6 + ((s-1 + s1)*4 + (s-2 + s2)); the border
//               column and the scalar tail use ((s0*6 + (s-1 + s1)*4) + s-2) + s2
//   vertical  : PyrDownVecV (4 lanes): ((r1 + r3) + r2)*4 + ((r0 + r4) + (r2 + r2));
//               scalar tail: ((r2*6 + (r1 + r3)*4) + r0) + r4
// One thread owns one output column and PD_RPT consecutive output rows: it forms the 2*PD_RPT+3 horizontally
// filtered input rows in registers (interior: three 8-byte loads per row, coalesced across the warp) and then the
// vertical taps.  No shared memory, no barrier.
// (pd_hform / pd_vform: ssk_prep.cuh)

template <int DEPTH>
__global__ void __launch_bounds__(256) k_pyrdown(const PyrDownArgs a) {
  constexpr int NR = 2 * PD_RPT + 3;
  const int b = blockIdx.z;
  Img src = a.src;
  if (a.src_ptrs) src.data = a.src_ptrs[b];
  float *dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int ox = blockIdx.x * PD_OW + (threadIdx.x & (PD_OW - 1));
  const int oy0 = blockIdx.y * PD_OH + (threadIdx.x / PD_OW) * PD_RPT;
  if (ox >= a.dst_cols || oy0 >= a.dst_rows) return;
  const int width0 = min((src.cols - 3) / 2 + 1, a.dst_cols);   // PyrDownInvoker: columns free of border handling
  const bool hsimd = ox >= 1 && ox < 1 + 4 * ((width0 - 1) / 4);
  const bool vsimd = ox < (a.dst_cols & ~3);
  const int ix = 2 * ox - 2, iy = 2 * oy0 - 2;
  float h[NR];
  const bool fast = DEPTH == SSK_32F && src.cn == 1 && ix >= 0 && ix + 5 < src.cols && iy >= 0 && iy + NR - 1 < src.rows &&
                    (src.step & 7) == 0 && (reinterpret_cast<uintptr_t>(src.data) & 7) == 0;
  if (fast) {
    const char *p = static_cast<const char *>(src.data) + (int64_t)iy * src.step + (int64_t)ix * 4;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float2 *q = reinterpret_cast<const float2 *>(p + (int64_t)r * src.step);
      const float2 A = __ldg(q), B = __ldg(q + 1);
      const float p4 = __ldg(reinterpret_cast<const float *>(q + 2));
      h[r] = pd_hform(A.x, A.y, B.x, B.y, p4, hsimd);
    }
  } else {
    const int border = a.border ? a.border : SSK_BORDER_REFLECT101;
    int xx[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) xx[c] = border_idx(ix + c, src.cols, border);
    // unrolled: the loads of all rows are in flight together (a rolled loop walks the rows one memory latency at a time,
    // which is the whole run time of the small pyramid levels)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int yy = border_idx(iy + r, src.rows, border);
      h[r] = pd_hform(load_gray<DEPTH>(src, yy, xx[0]), load_gray<DEPTH>(src, yy, xx[1]), load_gray<DEPTH>(src, yy, xx[2]),
                      load_gray<DEPTH>(src, yy, xx[3]), load_gray<DEPTH>(src, yy, xx[4]), hsimd);
    }
  }
#pragma unroll
  for (int j = 0; j < PD_RPT; ++j) {
    if (oy0 + j >= a.dst_rows) break;
    const float r0 = h[2 * j], r1 = h[2 * j + 1], r2 = h[2 * j + 2], r3 = h[2 * j + 3], r4 = h[2 * j + 4];
    float v = pd_vform(r0, r1, r2, r3, r4, vsimd);
    if (a.post_scale != 1.f) v = __fmul_rn(v, a.post_scale);
    dst[(int64_t)(oy0 + j) * a.dst_cols + ox] = v;
  }
}

// ---- cv::resize INTER_AREA, down-scaling ------------------------------------------------------------------------
// Arithmetic restated from OpenCV's resize.cpp and checked bit-exactly against cv2 4.13 on the CPU (numpy model) before
// it was written here:
//   integer scale (ResizeAreaFast): sum of the iscale_x * iscale_y cell in row-major order, four terms at a time
//     (s += ((a + b) + c) + d), times 1/area; for 2 x 2 the columns covered by the 8-lane SIMD loop use
//     ((s00 + s01) + (s10 + s11)) * 0.25 (the tail columns of a row use the scalar form; the lane count is the
//     AVX2 one - a build with another vector width differs in the last ulp on those few columns)
//   otherwise (ResizeArea): per destination pixel the source cells it overlaps, weights alpha / beta =
//     overlap / cellWidth computed in double and narrowed to float (computeResizeAreaTab); a row's contribution is
//     buf = sum_x S * alpha (in x order), the pixel is sum_y beta * buf (first row assigns, the others add).
constexpr int kAreaSimdLanes = 8;

struct AreaAxis { int s0, n; float a_first, a_mid, a_last; bool has_first, has_last; };

__device__ __forceinline__ AreaAxis area_axis(int d, double scale, int ssize) {
  AreaAxis r;
  const double fsx1 = d * scale;
  const double cell = fmin(scale, ssize - fsx1);
  const double fsx2 = fsx1 + cell;
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  r.has_first = sx1 - fsx1 > 1e-3;
  r.has_last = fsx2 - sx2 > 1e-3;
  r.a_first = (float)((sx1 - fsx1) / cell);
  r.a_mid = (float)(1.0 / cell);
  r.a_last = (float)(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
  r.s0 = r.has_first ? sx1 - 1 : sx1;
  r.n = (sx2 - sx1) + (r.has_first ? 1 : 0) + (r.has_last ? 1 : 0);
  return r;
}
__device__ __forceinline__ float area_weight(const AreaAxis &r, int k) {
  if (k == 0 && r.has_first) return r.a_first;
  if (k == r.n - 1 && r.has_last) return r.a_last;
  return r.a_mid;
}

// the per-axis tables of the fractional path, one entry per destination column / row (the double-precision divisions of
// area_axis once per axis instead of twice per pixel)
__global__ void __launch_bounds__(256) k_area_axis_tab(int dcols, int drows, double scale_x, double scale_y, int scols, int srows, AreaAxis *tab) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < dcols) tab[i] = area_axis(i, scale_x, scols);
  else if (i < dcols + drows) tab[i] = area_axis(i - dcols, scale_y, srows);
}

template <int DEPTH>
__global__ void __launch_bounds__(256) k_resize_area(const ResizeAreaArgs a, double scale_x, double scale_y, int iscale_x, int iscale_y,
                                                     const AreaAxis *__restrict__ axis_tab) {
  const int b = blockIdx.z;
  Img src = a.src;
  if (a.src_ptrs) src.data = a.src_ptrs[b];
  float *dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (dx >= a.dst_cols || dy >= a.dst_rows) return;
  float out;
  if (iscale_x > 0) {
    const int x0 = dx * iscale_x, y0 = dy * iscale_y;
    const bool full = x0 + iscale_x <= src.cols && y0 + iscale_y <= src.rows;
    const int wfull = min(a.dst_cols, src.cols / iscale_x);      // columns whose cell is complete (ResizeAreaFast_Invoker's w)
    if (!full) {
      // cell clipped by the image edge: plain running sum over the part inside, divided by the number of samples
      float sum = 0.f;
      int count = 0;
      for (int ky = 0; ky < iscale_y && y0 + ky < src.rows; ++ky)
        for (int kx = 0; kx < iscale_x && x0 + kx < src.cols; ++kx) { sum = __fadd_rn(sum, load_gray<DEPTH>(src, y0 + ky, x0 + kx)); ++count; }
      out = count ? __fdiv_rn(sum, (float)count) : 0.f;
    } else if (iscale_x == 2 && iscale_y == 2 && dx < wfull - wfull % kAreaSimdLanes) {
      const float s00 = load_gray<DEPTH>(src, y0, x0), s01 = load_gray<DEPTH>(src, y0, x0 + 1);
      const float s10 = load_gray<DEPTH>(src, y0 + 1, x0), s11 = load_gray<DEPTH>(src, y0 + 1, x0 + 1);
      out = __fmul_rn(__fadd_rn(__fadd_rn(s00, s01), __fadd_rn(s10, s11)), 0.25f);
    } else {
      const int area = iscale_x * iscale_y;
      float sum = 0.f;
      int k = 0;
      auto at = [&](int kk) { const int ky = kk / iscale_x, kx = kk - ky * iscale_x; return load_gray<DEPTH>(src, y0 + ky, x0 + kx); };
      for (; k + 4 <= area; k += 4) sum = __fadd_rn(sum, __fadd_rn(__fadd_rn(__fadd_rn(at(k), at(k + 1)), at(k + 2)), at(k + 3)));
      for (; k < area; ++k) sum = __fadd_rn(sum, at(k));
      out = __fmul_rn(sum, (float)(1.0 / area));
    }
  } else {
    const AreaAxis ax = axis_tab ? axis_tab[dx] : area_axis(dx, scale_x, src.cols);
    const AreaAxis ay = axis_tab ? axis_tab[a.dst_cols + dy] : area_axis(dy, scale_y, src.rows);
    out = 0.f;
    for (int ky = 0; ky < ay.n; ++ky) {
      float buf = 0.f;
      for (int kx = 0; kx < ax.n; ++kx) buf = __fadd_rn(buf, __fmul_rn(load_gray<DEPTH>(src, ay.s0 + ky, ax.s0 + kx), area_weight(ax, kx)));
      const float t = __fmul_rn(area_weight(ay, ky), buf);
      out = ky == 0 ? t : __fadd_rn(out, t);
    }
  }
  dst[(int64_t)dy * a.dst_cols + dx] = out;
}

template <int DEPTH>
__global__ void __launch_bounds__(256) k_to_gray(const Img im, const void *const *src_ptrs, float *dst, float *const *dst_ptrs) {
  const int b = blockIdx.z;
  Img src = im;
  if (src_ptrs) src.data = src_ptrs[b];
  float *d = dst_ptrs ? dst_ptrs[b] : dst;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x < src.cols && y < src.rows) d[(int64_t)y * src.cols + x] = load_gray<DEPTH>(src, y, x);
}

constexpr int SF_W = 64, SF_H = 16, SF_R = (kMaxTaps - 1) / 2;

// KXN / KYN > 0: tap counts known at compile time (loops unroll); 0: taken from the arguments
template <int KXN, int KYN>
__global__ void __launch_bounds__(256) k_sepfilter(const SepFilterArgs a) {
  constexpr int RMAX = (KXN && KYN) ? ((KXN > KYN ? KXN : KYN) >> 1) : SF_R;
  // the per-frame Gaussians (7 taps: the ECC image smoothing, 9 taps: cv::GaussianBlur(sigma = 1) of the weight maps):
  // register-blocked passes
  constexpr bool G7 = KXN == KYN && (KXN == 7 || KXN == 9);
  constexpr int PITCH = G7 ? SF_W + 8 : SF_W + 2 * RMAX + 1;     // G7: a multiple of 4 floats (16-byte shared loads)
  __shared__ __align__(16) float s_in[SF_H + 2 * RMAX][PITCH];
  __shared__ __align__(16) float s_h[SF_H + 2 * RMAX][SF_W];
  // interleaved channels and borders other than REPLICATE go through the generic instantiation only
  const int cn = (KXN || a.cn < 1) ? 1 : a.cn;
  const int b = blockIdx.z / cn, ch = blockIdx.z - b * cn;
  const float *src = a.src_ptrs ? a.src_ptrs[b] : a.src;
  float *dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int kxn = KXN ? KXN : a.kxn, kyn = KYN ? KYN : a.kyn;
  const int rx = kxn >> 1, ry = kyn >> 1;
  const int x0 = blockIdx.x * SF_W, y0 = blockIdx.y * SF_H;
  const int ow = min(SF_W, a.cols - x0), oh = min(SF_H, a.rows - y0);
  const int iw = ow + 2 * rx, ih = oh + 2 * ry;
  // register-blocked Gaussian passes (below): the staged tile starts at column x0 - 4 (16-byte groups of the source row), so
  // the sample of column x0 - rx sits at index 4 - rx
  const bool g7 = G7 && a.ky[0] == a.ky[(KYN ? KYN : 1) - 1];
  const int off = g7 ? 4 - rx : 0;
  if (g7 && x0 >= 4 && x0 + SF_W + 4 <= a.cols && (a.cols & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // tile away from the left / right border: 18 aligned 16-byte loads per row instead of 70 scalar ones
    for (int i = threadIdx.x; i < ih * (PITCH / 4); i += 256) {
      const int r = i / (PITCH / 4), q = i - r * (PITCH / 4);
      const int yy = min(max(y0 - ry + r, 0), a.rows - 1);
      *reinterpret_cast<float4 *>(&s_in[r][4 * q]) = __ldg(reinterpret_cast<const float4 *>(src + (int64_t)yy * a.cols + x0 - 4) + q);
    }
  } else {
    for_tile(ih, iw, [&](int r, int c) {
      int yy, xx;
      if (KXN == 0 && a.border) {
        yy = border_idx(y0 - ry + r, a.rows, a.border); xx = border_idx(x0 - rx + c, a.cols, a.border);
      } else {
        yy = min(max(y0 - ry + r, 0), a.rows - 1); xx = min(max(x0 - rx + c, 0), a.cols - 1);
      }
      s_in[r][c + off] = __ldg(src + ((int64_t)yy * a.cols + xx) * cn + ch);
    });
  }
  __syncthreads();
  // row / column arithmetic follows OpenCV's filter engine (found bit-exact against cv2 4.13 for the kernels of this
  // path: 5-tap derivative, 3-tap smoothing, 7-tap Gaussian)
  if constexpr (G7) {
    if (g7) {
      // Same operations in the same order as the generic passes below, four outputs per thread: the row pass reads its
      // (3 + taps)-sample window with three 16-byte shared loads (instead of 4 x taps scalar ones), the column pass keeps a
      // (3 + taps)-row window in registers.
      constexpr int R = KXN >> 1;
      for (int i = threadIdx.x; i < (SF_H + 2 * R) * (SF_W / 4); i += 256) {
        const int r = i / (SF_W / 4), cg = (i % (SF_W / 4)) * 4;
        if (r >= ih) break;
        const float4 v0 = *reinterpret_cast<const float4 *>(&s_in[r][cg]);
        const float4 v1 = *reinterpret_cast<const float4 *>(&s_in[r][cg + 4]);
        const float4 v2 = *reinterpret_cast<const float4 *>(&s_in[r][cg + 8]);
        const float w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          constexpr int OFF = 4 - R;                      // index of column x0 - R in the staged row
          float acc = __fmul_rn(w[j + OFF], a.kx[0]);    // RowVec_32f: taps in order, fma chain
#pragma unroll
          for (int t = 1; t < KXN; ++t) acc = __fmaf_rn(w[j + t + OFF], a.kx[t], acc);
          o[j] = acc;
        }
        *reinterpret_cast<float4 *>(&s_h[r][cg]) = make_float4(o[0], o[1], o[2], o[3]);
      }
      __syncthreads();
      const int tx = threadIdx.x & (SF_W - 1), g4 = (threadIdx.x / SF_W) * 4;
      float c[4 + 2 * R];
#pragma unroll
      for (int q = 0; q < 4 + 2 * R; ++q) c[q] = s_h[g4 + q][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = __fmul_rn(a.ky[R], c[j + R]);        // SymmColumnVec_32f: centre tap, then fma over the pairs
#pragma unroll
        for (int t = 1; t <= R; ++t) acc = __fmaf_rn(__fadd_rn(c[j + R + t], c[j + R - t]), a.ky[R + t], acc);
        const int ty = g4 + j;
        if (ty < oh && tx < ow) dst[(int64_t)(y0 + ty) * a.cols + x0 + tx] = acc;
      }
      return;
    }
  }
  const bool xsym = a.kx[0] == a.kx[kxn - 1];
  for_tile(ih, ow, [&](int r, int x) {
    const float *p = &s_in[r][x + rx];
    float acc;
    if (kxn <= 5) {
      // SymmRowSmallFilter (SymmRowSmallVec_32f): symmetric 3 taps fma(x0, k0, (x-1 + x1) * k1), 5 taps
      // fma(x-2 + x2, k2, that) (found against cv2 4.13 with a general Gaussian, tests/test_unsharp_oracle.py; for the
      // power-of-two smoothing kernel of ecc_differentiate every product is exact and the order is immaterial);
      // antisymmetric: first pair product, then fma over the remaining pairs
      if (xsym) {
        if (rx == 0) {
          acc = __fmul_rn(a.kx[0], p[0]);
        } else {
          acc = __fmaf_rn(p[0], a.kx[rx], __fmul_rn(__fadd_rn(p[1], p[-1]), a.kx[rx + 1]));
          if (rx == 2) acc = __fmaf_rn(__fadd_rn(p[2], p[-2]), a.kx[rx + 2], acc);
        }
      } else {
        acc = rx >= 1 ? __fmul_rn(__fsub_rn(p[1], p[-1]), a.kx[rx + 1]) : 0.f;
#pragma unroll
        for (int i = 2; i <= rx; ++i) acc = __fmaf_rn(__fsub_rn(p[i], p[-i]), a.kx[rx + i], acc);
      }
    } else {
      // RowFilter (RowVec_32f): taps in order, fma chain
      acc = __fmul_rn(p[-rx], a.kx[0]);
#pragma unroll
      for (int i = 1; i < kxn; ++i) acc = __fmaf_rn(p[i - rx], a.kx[i], acc);
    }
    s_h[r][x] = acc;
  });
  __syncthreads();
  const bool ysym = a.ky[0] == a.ky[kyn - 1];
  for_tile(oh, ow, [&](int ty, int tx) {
    const int r = ty + ry;
    // SymmColumnFilter (SymmColumnVec_32f): centre tap then fma over the (anti)symmetric pairs
    float acc;
    if (ysym) {
      acc = __fmul_rn(a.ky[ry], s_h[r][tx]);
#pragma unroll
      for (int i = 1; i <= ry; ++i) acc = __fmaf_rn(__fadd_rn(s_h[r + i][tx], s_h[r - i][tx]), a.ky[ry + i], acc);
    } else {
      acc = ry >= 1 ? __fmul_rn(__fsub_rn(s_h[r + 1][tx], s_h[r - 1][tx]), a.ky[ry + 1]) : 0.f;
#pragma unroll
      for (int i = 2; i <= ry; ++i) acc = __fmaf_rn(__fsub_rn(s_h[r + i][tx], s_h[r - i][tx]), a.ky[ry + i], acc);
    }
    dst[((int64_t)(y0 + ty) * a.cols + x0 + tx) * cn + ch] = acc;
  });
}

// ---- W1 ------------------------------------------------------------------------------------------
constexpr int W1_W = 64, W1_H = 32, W1_RMAX = 4;

__global__ void __launch_bounds__(256) k_w1_grad(const W1Args a, int nblocks) {
  __shared__ float s_in[W1_H + 2 * W1_RMAX][W1_W + 2 * W1_RMAX + 1];
  __shared__ double s_red[2][8];
  const int b = blockIdx.z;
  const float *M = a.M_ptrs ? a.M_ptrs[b] : a.M;
  float *gmap = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  const int r = a.kradius;
  const int x0 = blockIdx.x * W1_W, y0 = blockIdx.y * W1_H;
  const int ow = min(W1_W, a.cols - x0), oh = min(W1_H, a.rows - y0);
  for_tile(oh + 2 * r, ow + 2 * r, [&](int rr, int c) {
    const int yy = min(max(y0 - r + rr, 0), a.rows - 1), xx = min(max(x0 - r + c, 0), a.cols - 1);
    s_in[rr][c] = __ldg(M + (int64_t)yy * a.cols + xx);
  });
  __syncthreads();
  const float ms = (float)(a.depth_scale * a.depth_scale * a.depth_scale);
  double sg = 0.0, sg4 = 0.0;
  for_tile(oh, ow, [&](int ty, int tx) {
    float mx = -3.4e38f, mn = 3.4e38f;
    for (int dy = 0; dy <= 2 * r; ++dy)
      for (int dx = 0; dx <= 2 * r; ++dx) {
        const float v = s_in[ty + dy][tx + dx];
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
      }
    const float g = __fsub_rn(mx, mn);
    gmap[(int64_t)(y0 + ty) * a.cols + x0 + tx] = __fmul_rn(__fmul_rn(__fmul_rn(g, g), g), ms);
    sg += (double)fabsf(g);
    sg4 += (double)__fmul_rn(__fmul_rn(__fmul_rn(g, g), g), g);
  });
  sg = warp_sum(sg);
  sg4 = warp_sum(sg4);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][warp] = sg; s_red[1][warp] = sg4; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int i = 0; i < 8; ++i) { t0 += s_red[0][i]; t1 += s_red[1][i]; }
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[((int64_t)b * 2 + 0) * nblocks + blk] = t0;
    a.partials[((int64_t)b * 2 + 1) * nblocks + blk] = t1;
  }
}

// kradius = 1 (the default): one thread owns one column and W1_H / 4 consecutive rows of the 64 x 16 tile, keeps the
// horizontal 3-tap max / min of its rows in registers (neighbours by warp shuffle) and combines them vertically.
// No shared-memory tile; the per-block partial sums keep the layout of k_w1_grad.
__global__ void __launch_bounds__(256) k_w1_grad_r1(const W1Args a, int nblocks) {
  constexpr int RPT = W1_H / 4;
  __shared__ double s_red[2][8];
  const int b = blockIdx.z;
  const float *__restrict__ M = a.M_ptrs ? a.M_ptrs[b] : a.M;
  float *__restrict__ gmap = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * W1_W + (threadIdx.x & (W1_W - 1));
  const int y0 = blockIdx.y * W1_H + (threadIdx.x / W1_W) * RPT;
  const int xc = min(x, a.cols - 1), xl = max(xc - 1, 0), xr = min(xc + 1, a.cols - 1);
  float hmax[RPT + 2], hmin[RPT + 2];
#pragma unroll
  for (int r = 0; r < RPT + 2; ++r) {
    const int yy = min(max(y0 - 1 + r, 0), a.rows - 1);
    const float *row = M + (int64_t)yy * a.cols;
    const float c = __ldg(row + xc);
    float l = __shfl_up_sync(0xffffffffu, c, 1), rt = __shfl_down_sync(0xffffffffu, c, 1);
    if (lane == 0 || x >= a.cols) l = __ldg(row + xl);          // columns past the image edge replicate the last one
    if (lane == 31 || x + 1 >= a.cols) rt = __ldg(row + xr);
    hmax[r] = fmaxf(fmaxf(l, c), rt);
    hmin[r] = fminf(fminf(l, c), rt);
  }
  const float ms = (float)(a.depth_scale * a.depth_scale * a.depth_scale);
  double sg = 0.0, sg4 = 0.0;
  if (x < a.cols) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (y0 + j >= a.rows) break;
      const float mx = fmaxf(fmaxf(hmax[j], hmax[j + 1]), hmax[j + 2]), mn = fminf(fminf(hmin[j], hmin[j + 1]), hmin[j + 2]);
      const float g = __fsub_rn(mx, mn);
      gmap[(int64_t)(y0 + j) * a.cols + x] = __fmul_rn(__fmul_rn(__fmul_rn(g, g), g), ms);
      sg += (double)fabsf(g);
      sg4 += (double)__fmul_rn(__fmul_rn(__fmul_rn(g, g), g), g);
    }
  }
  sg = warp_sum(sg);
  sg4 = warp_sum(sg4);
  if (lane == 0) { s_red[0][warp] = sg; s_red[1][warp] = sg4; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0, t1 = 0;
    for (int i = 0; i < 8; ++i) { t0 += s_red[0][i]; t1 += s_red[1][i]; }
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[((int64_t)b * 2 + 0) * nblocks + blk] = t0;
    a.partials[((int64_t)b * 2 + 1) * nblocks + blk] = t1;
  }
}

__global__ void __launch_bounds__(256) k_w1_final(const W1Args a, int nblocks) {
  __shared__ double s0[256], s1[256];
  const int b = blockIdx.x;
  double t0 = 0, t1 = 0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    t0 += a.partials[((int64_t)b * 2 + 0) * nblocks + i];
    t1 += a.partials[((int64_t)b * 2 + 1) * nblocks + i];
  }
  s0[threadIdx.x] = t0; s1[threadIdx.x] = t1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double W = s0[0];
    const double ds3 = a.depth_scale * a.depth_scale * a.depth_scale;
    // the reference accumulates g^4 in a float (atomic CAS loop); narrow the total once
    const double Q = W > 0 ? (ds3 / W) * (double)(float)s1[0] : 0.0;
    a.stats[b * 4 + 0] = W;
    a.stats[b * 4 + 1] = s1[0];
    a.stats[b * 4 + 2] = Q;
    a.stats[b * 4 + 3] = (double)(float)(0.05 * Q);
  }
}

// cv::resize(map + 0.05 Q, full size, INTER_LINEAR) (c_local_variance_sharpness_measure.cc:176-184, 239-243).
// The source index and fraction of an output column / row do not depend on the frame: a tiny kernel tabulates them
// (fx = (float)((dx + 0.5) * scale - 0.5) in double, clamped like cv::resize), the up-sampling kernel reads them.
// Layout: idx[npad] then frac[npad] (npad = n_full rounded up to 4), so that a thread's 4 entries are one 16-byte load.
__global__ void __launch_bounds__(256) k_w1_axis(int n_full, int n_small, double sc, int *idx, float *frac) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_full) return;
  float f = (float)((i + 0.5) * sc - 0.5);
  int s = (int)floorf(f);
  f -= s;
  if (s < 0) { f = 0; s = 0; }
  if (s >= n_small - 1) { f = 0; s = n_small - 1; }
  idx[i] = s;
  frac[i] = f;
}

// One thread per 4 consecutive output pixels (one 16-byte store when the row allows it).
__global__ void __launch_bounds__(256) k_w1_upsample(const W1Args a) {
  const int b = blockIdx.z;
  const float *__restrict__ g = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  float *__restrict__ out = a.out_ptrs ? a.out_ptrs[b] : a.out;
  const int x4 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x4 >= a.full_cols || y >= a.full_rows) return;
  const float add = (float)a.stats[b * 4 + 3];
  float v[4];
  if (a.full_cols == a.cols && a.full_rows == a.rows) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __fadd_rn(__ldg(g + (int64_t)y * a.cols + min(x4 + k, a.cols - 1)), add);
  } else {
    const int xpad = (a.full_cols + 3) & ~3, ypad = (a.full_rows + 3) & ~3;
    const int *__restrict__ xi = reinterpret_cast<const int *>(a.axis_tab);
    const float *__restrict__ xf = reinterpret_cast<const float *>(xi + xpad);
    const int *__restrict__ yi = xi + 2 * xpad;
    const float *__restrict__ yf = reinterpret_cast<const float *>(yi + ypad);
    const int sy = __ldg(yi + y), sy1 = min(sy + 1, a.rows - 1);
    const float b1 = __ldg(yf + y), b0 = 1.f - b1;
    const int4 sx4 = __ldg(reinterpret_cast<const int4 *>(xi + x4));        // x4 is a multiple of 4, tables are padded
    const float4 fx4 = __ldg(reinterpret_cast<const float4 *>(xf + x4));
    const int sxs[4] = {sx4.x, sx4.y, sx4.z, sx4.w};
    const float fxs[4] = {fx4.x, fx4.y, fx4.z, fx4.w};
    const float *__restrict__ g0 = g + sy * a.cols, *__restrict__ g1 = g + sy1 * a.cols;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int sx = min(max(sxs[k], 0), a.cols - 1), sx1 = min(sx + 1, a.cols - 1);   // padding entries are clamped
      const float a1 = fxs[k], a0 = 1.f - a1;
      const float v00 = __fadd_rn(__ldg(g0 + sx), add), v01 = __fadd_rn(__ldg(g0 + sx1), add);
      const float v10 = __fadd_rn(__ldg(g1 + sx), add), v11 = __fadd_rn(__ldg(g1 + sx1), add);
      const float r0 = __fadd_rn(__fmul_rn(v00, a0), __fmul_rn(v01, a1));
      const float r1 = __fadd_rn(__fmul_rn(v10, a0), __fmul_rn(v11, a1));
      v[k] = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
    }
  }
  float *o = out + (int64_t)y * a.full_cols + x4;
  if (x4 + 3 < a.full_cols && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
    *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int k = 0; k < 4 && x4 + k < a.full_cols; ++k) o[k] = v[k];
  }
}

// Exact 2x up-sampling (full = 2 x small in both axes, the default dscale = 1): one thread produces a 4 x 2 block of
// outputs from a 4 x 3 window of the small map (1.5 loads per output instead of 4).  Same tables, same operand order
// as k_w1_upsample; threads whose window is clipped by the image edge take the table-driven path element by element.
__device__ __forceinline__ float w1_bilin(float v00, float v01, float v10, float v11, float a1, float b1) {
  const float a0 = 1.f - a1, b0 = 1.f - b1;
  const float r0 = __fadd_rn(__fmul_rn(v00, a0), __fmul_rn(v01, a1));
  const float r1 = __fadd_rn(__fmul_rn(v10, a0), __fmul_rn(v11, a1));
  return __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
}

__global__ void __launch_bounds__(256) k_w1_upsample2x(const W1Args a) {
  const int b = blockIdx.z;
  const float *__restrict__ g = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  float *__restrict__ out = a.out_ptrs ? a.out_ptrs[b] : a.out;
  const int x4 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, y0 = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 2;
  if (x4 >= a.full_cols || y0 >= a.full_rows) return;
  const float add = (float)a.stats[b * 4 + 3];
  const int xpad = (a.full_cols + 3) & ~3, ypad = (a.full_rows + 3) & ~3;
  const int *__restrict__ xi = reinterpret_cast<const int *>(a.axis_tab);
  const float *__restrict__ xf = reinterpret_cast<const float *>(xi + xpad);
  const int *__restrict__ yi = xi + 2 * xpad;
  const float *__restrict__ yf = reinterpret_cast<const float *>(yi + ypad);
  const int4 sx4 = __ldg(reinterpret_cast<const int4 *>(xi + x4));
  const float4 fx4 = __ldg(reinterpret_cast<const float4 *>(xf + x4));
  const int2 sy2 = __ldg(reinterpret_cast<const int2 *>(yi + y0));          // y0 is even, tables are padded
  const float2 fy2 = __ldg(reinterpret_cast<const float2 *>(yf + y0));
  const int sxs[4] = {sx4.x, sx4.y, sx4.z, sx4.w};
  const float fxs[4] = {fx4.x, fx4.y, fx4.z, fx4.w};
  const int sys[2] = {sy2.x, sy2.y};
  const float fys[2] = {fy2.x, fy2.y};
  float v[2][4];
  const int base = sx4.x;
  const bool regular = x4 + 3 < a.full_cols && y0 + 1 < a.full_rows && base >= 0 && base + 3 < a.cols && sx4.y == base + 1 &&
                       sx4.z == base + 1 && sx4.w == base + 2 && sy2.x >= 0 && sy2.y == sy2.x + 1 && sy2.y + 1 < a.rows;
  if (regular) {
    float w[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float *__restrict__ row = g + (sy2.x + r) * a.cols + base;
#pragma unroll
      for (int q = 0; q < 4; ++q) w[r][q] = __fadd_rn(__ldg(row + q), add);
    }
    // columns: offsets {0, 1, 1, 2} into the window; rows: {0, 1} for y0 and {1, 2} for y0 + 1
    v[0][0] = w1_bilin(w[0][0], w[0][1], w[1][0], w[1][1], fxs[0], fys[0]);
    v[0][1] = w1_bilin(w[0][1], w[0][2], w[1][1], w[1][2], fxs[1], fys[0]);
    v[0][2] = w1_bilin(w[0][1], w[0][2], w[1][1], w[1][2], fxs[2], fys[0]);
    v[0][3] = w1_bilin(w[0][2], w[0][3], w[1][2], w[1][3], fxs[3], fys[0]);
    v[1][0] = w1_bilin(w[1][0], w[1][1], w[2][0], w[2][1], fxs[0], fys[1]);
    v[1][1] = w1_bilin(w[1][1], w[1][2], w[2][1], w[2][2], fxs[1], fys[1]);
    v[1][2] = w1_bilin(w[1][1], w[1][2], w[2][1], w[2][2], fxs[2], fys[1]);
    v[1][3] = w1_bilin(w[1][2], w[1][3], w[2][2], w[2][3], fxs[3], fys[1]);
  } else {
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      if (y0 + j >= a.full_rows) break;
      const int sy = sys[j], sy1 = min(sy + 1, a.rows - 1);
      const float *__restrict__ g0 = g + sy * a.cols, *__restrict__ g1 = g + sy1 * a.cols;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int sx = min(max(sxs[k], 0), a.cols - 1), sx1 = min(sx + 1, a.cols - 1);   // padding entries are clamped
        v[j][k] = w1_bilin(__fadd_rn(__ldg(g0 + sx), add), __fadd_rn(__ldg(g0 + sx1), add), __fadd_rn(__ldg(g1 + sx), add),
                           __fadd_rn(__ldg(g1 + sx1), add), fxs[k], fys[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (y0 + j >= a.full_rows) break;
    float *o = out + (int64_t)(y0 + j) * a.full_cols + x4;
    if (x4 + 3 < a.full_cols && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      *reinterpret_cast<float4 *>(o) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
    } else {
      for (int k = 0; k < 4 && x4 + k < a.full_cols; ++k) o[k] = v[j][k];
    }
  }
}

// The same up-sampling, eight outputs x two rows per thread (the form the stacking loop runs: 1080p from a 960 x 540 map).
// For full = 2 x small the table of k_w1_axis is s = k - 1, f = 0.75 for dx = 2k and s = k, f = 0.25 for dx = 2k + 1
// ((dx + 0.5) * 0.5 - 0.5 is exact); only the first and the last sample of an axis are clamped (s = 0 resp. n - 1 with
// f = 0, the second tap being s + 1 clamped to n - 1).  So a thread needs no table: it reads one aligned 16-byte group of
// three map rows, takes the two neighbouring samples from the lanes beside it, forms the horizontal interpolations of the
// three rows once (cv::resize's order: v0 * (1 - f) + v1 * f, each product rounded) and combines them vertically; the
// threads on the image border substitute the clamped taps with f = 0 (evaluated literally: v0 * 1 + v1 * 0).  Only frames
// whose buffers are not 16-byte aligned take the table-driven path element by element.  (First version: border threads
// walked the tables; one such lane per warp on the left / right border made a quarter of the warps 10 x slower.)
__global__ void __launch_bounds__(256) k_w1_upsample2x_v8(const W1Args a) {
  const int b = blockIdx.z;
  const float *__restrict__ g = a.gmap_ptrs ? a.gmap_ptrs[b] : a.gmap;
  float *__restrict__ out = a.out_ptrs ? a.out_ptrs[b] : a.out;
  const int lane = threadIdx.x & 31;
  const int j4 = (blockIdx.x * 32 + lane) * 4;                   // first map column of this thread's group
  const int i = blockIdx.y * 8 + (threadIdx.x >> 5);             // map row: output rows 2i and 2i + 1
  const int x8 = 2 * j4, y0 = 2 * i;
  const bool inside = j4 < a.cols && i < a.rows;                 // cols is a multiple of 4: a group is inside as a whole
  const float add = (float)a.stats[b * 4 + 3];
  const bool aligned = ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;   // 16-byte loads and stores
  const bool fast = inside && aligned;
  const bool first = j4 == 0, last = j4 + 4 == a.cols, top = i == 0, bottom = i == a.rows - 1;
  const unsigned fast_mask = __ballot_sync(0xffffffffu, fast);
  float w[3][6];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *__restrict__ row = g + (int64_t)min(max(i - 1 + r, 0), a.rows - 1) * a.cols;
    if (fast) v = __ldg(reinterpret_cast<const float4 *>(row + j4));
    float left = __shfl_up_sync(0xffffffffu, v.w, 1), right = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (fast) {
      // a neighbour that did not load (another warp, or outside the map): fetch the sample directly; none beyond the border
      if (first) left = 0.f;
      else if (lane == 0 || !((fast_mask >> (lane - 1)) & 1u)) left = __ldg(row + j4 - 1);
      if (last) right = 0.f;
      else if (lane == 31 || !((fast_mask >> (lane + 1)) & 1u)) right = __ldg(row + j4 + 4);
    }
    w[r][0] = __fadd_rn(left, add); w[r][1] = __fadd_rn(v.x, add); w[r][2] = __fadd_rn(v.y, add);
    w[r][3] = __fadd_rn(v.z, add); w[r][4] = __fadd_rn(v.w, add); w[r][5] = __fadd_rn(right, add);
  }
  if (!inside) return;
  if (fast) {
    float h[3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        // t even: samples (t / 2, t / 2 + 1) of the window with f = 0.75; t odd: ((t + 1) / 2, (t + 1) / 2 + 1) with f = 0.25
        const int k = (t + 1) >> 1;
        const float a1 = (t & 1) ? 0.25f : 0.75f, a0 = 1.f - a1;
        h[r][t] = __fadd_rn(__fmul_rn(w[r][k], a0), __fmul_rn(w[r][k + 1], a1));
      }
      // dx = 0: s = 0, f = 0 (taps 0 and 1 of the map); dx = 2 cols - 1: s = cols - 1, f = 0 (second tap clamped onto the first)
      if (first) h[r][0] = __fadd_rn(__fmul_rn(w[r][1], 1.f), __fmul_rn(w[r][2], 0.f));
      if (last) h[r][7] = __fadd_rn(__fmul_rn(w[r][4], 1.f), __fmul_rn(w[r][4], 0.f));
    }
    float o0[8], o1[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      // output row 2i: s = i - 1, f = 0.75 (top row: s = 0, f = 0, taps rows 0 and 1); row 2i + 1: s = i, f = 0.25 (bottom row:
      // s = rows - 1, f = 0, second tap clamped onto the first)
      o0[t] = top ? __fadd_rn(__fmul_rn(h[1][t], 1.f), __fmul_rn(h[2][t], 0.f)) : __fadd_rn(__fmul_rn(h[0][t], 0.25f), __fmul_rn(h[1][t], 0.75f));
      o1[t] = bottom ? __fadd_rn(__fmul_rn(h[1][t], 1.f), __fmul_rn(h[1][t], 0.f)) : __fadd_rn(__fmul_rn(h[1][t], 0.75f), __fmul_rn(h[2][t], 0.25f));
    }
    float4 *p0 = reinterpret_cast<float4 *>(out + (int64_t)y0 * a.full_cols + x8);
    float4 *p1 = reinterpret_cast<float4 *>(out + (int64_t)(y0 + 1) * a.full_cols + x8);
    p0[0] = make_float4(o0[0], o0[1], o0[2], o0[3]); p0[1] = make_float4(o0[4], o0[5], o0[6], o0[7]);
    p1[0] = make_float4(o1[0], o1[1], o1[2], o1[3]); p1[1] = make_float4(o1[4], o1[5], o1[6], o1[7]);
    return;
  }
  // buffers that are not 16-byte aligned: the tables, one element at a time
  const int xpad = (a.full_cols + 3) & ~3, ypad = (a.full_rows + 3) & ~3;
  const int *__restrict__ xi = reinterpret_cast<const int *>(a.axis_tab);
  const float *__restrict__ xf = reinterpret_cast<const float *>(xi + xpad);
  const int *__restrict__ yi = xi + 2 * xpad;
  const float *__restrict__ yf = reinterpret_cast<const float *>(yi + ypad);
#pragma unroll 1
  for (int jj = 0; jj < 2; ++jj) {
    const int y = y0 + jj;
    if (y >= a.full_rows) break;
    const int sy = __ldg(yi + y), sy1 = min(sy + 1, a.rows - 1);
    const float b1 = __ldg(yf + y);
    const float *__restrict__ g0 = g + (int64_t)sy * a.cols, *__restrict__ g1 = g + (int64_t)sy1 * a.cols;
#pragma unroll 1
    for (int t = 0; t < 8; ++t) {
      const int x = x8 + t;
      if (x >= a.full_cols) break;
      const int sx = min(max(__ldg(xi + x), 0), a.cols - 1), sx1 = min(sx + 1, a.cols - 1);
      out[(int64_t)y * a.full_cols + x] = w1_bilin(__fadd_rn(__ldg(g0 + sx), add), __fadd_rn(__ldg(g0 + sx1), add), __fadd_rn(__ldg(g1 + sx), add),
                                                  __fadd_rn(__ldg(g1 + sx1), add), __ldg(xf + x), b1);
    }
  }
}

__global__ void __launch_bounds__(256) k_erode5_u8(const uint8_t *src, int64_t sstep, uint8_t *dst, int64_t dstep, int rows,
                                                   int cols, int border_replicate) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  uint8_t m = 255;
  for (int dy = -2; dy <= 2; ++dy)
    for (int dx = -2; dx <= 2; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (border_replicate) { yy = min(max(yy, 0), rows - 1); xx = min(max(xx, 0), cols - 1); }
      else if ((unsigned)yy >= (unsigned)rows || (unsigned)xx >= (unsigned)cols) continue;
      m = min(m, src[(int64_t)yy * sstep + xx]);
    }
  dst[(int64_t)y * dstep + x] = m;
}


// ---- pyrUp ---------------------------------------------------------------------------------------
// cv::pyrUp on CV_32FC1, bit-exact against cv2 4.13 (tests/test_cvmodel.py::test_pyrup_model):
//   row pass : even column 2x = (s[x-1] + 6 s[x]) + s[x+1]   (x = 0: 6 s0 + 2 s1;  x = w-1: s[w-2] + 7 s[w-1])
//              odd  column 2x+1 = 4 (s[x] + s[x+1])            (x = w-1: 8 s[w-1]); an extra last column repeats it
//   col pass : even row 2y = (6 r[y] + r[y-1]) + r[y+1], odd row = 4 (r[y] + r[y+1]) with r[-1] = r[1], r[h] = r[h-1];
//              result * 1/64
__device__ __forceinline__ float pu_row(const float *__restrict__ p, int w, int ox) {
  const int x = min(ox >> 1, w - 1);
  const bool odd = (ox & 1) || (ox >> 1) >= w;
  if (odd) return x == w - 1 ? __fmul_rn(p[x], 8.f) : __fmul_rn(__fadd_rn(p[x], p[x + 1]), 4.f);
  if (x == 0) return __fadd_rn(__fmul_rn(p[0], 6.f), __fmul_rn(p[1], 2.f));
  if (x == w - 1) return __fadd_rn(p[x - 1], __fmul_rn(p[x], 7.f));
  return __fadd_rn(__fadd_rn(p[x - 1], __fmul_rn(p[x], 6.f)), p[x + 1]);
}

// one output of cv::pyrUp (any position, all border rules) and the optional `minuend - pyrUp` epilogue of ecc_normalize
__device__ __forceinline__ float pu_px(const float *__restrict__ src, int w, int h, int ox, int oy) {
  const int y = min(oy >> 1, h - 1);
  const bool odd = (oy & 1) || (oy >> 1) >= h;
  const int yd = min(y + 1, h - 1), yu = y == 0 ? 1 : y - 1;
  const float r1 = pu_row(src + (int64_t)y * w, w, ox), r2 = pu_row(src + (int64_t)yd * w, w, ox);
  float v;
  if (odd) v = __fmul_rn(__fadd_rn(r1, r2), 4.f);
  else v = __fadd_rn(__fadd_rn(__fmul_rn(r1, 6.f), pu_row(src + (int64_t)yu * w, w, ox)), r2);
  return __fmul_rn(v, 1.0f / 64.0f);
}

__device__ __forceinline__ float pu_epilogue(const PyrUpArgs &a, int b, int64_t o, float v) {
  if (a.minuend_ptrs || a.minuend) {
    const float *m = a.minuend_ptrs ? a.minuend_ptrs[b] : a.minuend;
    v = __fsub_rn(m[o], v);
    if (a.mask && a.mask[o] == 0) v = 0.f;
  }
  return v;
}

__global__ void __launch_bounds__(256) k_pyrup(const PyrUpArgs a) {
  const int b = blockIdx.z;
  const float *__restrict__ src = a.src_ptrs ? a.src_ptrs[b] : a.src;
  float *__restrict__ dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= a.dst_cols || oy >= a.dst_rows) return;
  const int64_t o = (int64_t)oy * a.dst_cols + ox;
  dst[o] = pu_epilogue(a, b, o, pu_px(src, a.cols, a.rows, ox, oy));
}

// The same operator, one thread per SOURCE pixel (x, y) -> the 2 x 2 outputs (2x .. 2x + 1, 2y .. 2y + 1).  Away from the
// borders the thread loads its own column of the rows y - 1, y, y + 1, takes the columns x - 1 and x + 1 from the lanes
// beside it, forms the even / odd row-pass values of the three rows once and combines them into the four outputs (the
// operations and their order are those of pu_row / pu_px); the two outputs of a row leave as one 8-byte store.  Border
// threads evaluate pu_px per output.  9 loads + ~70 instructions per output become 3 loads + 6 shuffles per four outputs.
__global__ void __launch_bounds__(256) k_pyrup_2x2(const PyrUpArgs a) {
  const int b = blockIdx.z;
  const float *__restrict__ src = a.src_ptrs ? a.src_ptrs[b] : a.src;
  float *__restrict__ dst = a.dst_ptrs ? a.dst_ptrs[b] : a.dst;
  const int lane = threadIdx.x & 31;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int w = a.cols, h = a.rows;
  const bool inside = 2 * x < a.dst_cols && 2 * y < a.dst_rows;
  const bool fast = inside && x >= 1 && x <= w - 2 && y >= 1 && y <= h - 2 && 2 * x + 1 < a.dst_cols && 2 * y + 1 < a.dst_rows &&
                    !(a.dst_cols & 1) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
  const unsigned fast_mask = __ballot_sync(0xffffffffu, fast);
  float c[3], l[3], r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float *__restrict__ row = src + (int64_t)min(max(y - 1 + k, 0), h - 1) * w;
    c[k] = fast ? __ldg(row + x) : 0.f;
    l[k] = __shfl_up_sync(0xffffffffu, c[k], 1);
    r[k] = __shfl_down_sync(0xffffffffu, c[k], 1);
    if (fast) {
      if (lane == 0 || !((fast_mask >> (lane - 1)) & 1u)) l[k] = __ldg(row + x - 1);
      if (lane == 31 || !((fast_mask >> (lane + 1)) & 1u)) r[k] = __ldg(row + x + 1);
    }
  }
  if (!inside) return;
  if (fast) {
    float E[3], O[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      E[k] = __fadd_rn(__fadd_rn(l[k], __fmul_rn(c[k], 6.f)), r[k]);      // pu_row, even column, interior
      O[k] = __fmul_rn(__fadd_rn(c[k], r[k]), 4.f);                       // pu_row, odd column, interior
    }
    // even output row 2y: (6 r[y] + r[y-1]) + r[y+1]; odd output row 2y + 1: 4 (r[y] + r[y+1])
    float v00 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(E[1], 6.f), E[0]), E[2]), 1.0f / 64.0f);
    float v01 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(O[1], 6.f), O[0]), O[2]), 1.0f / 64.0f);
    float v10 = __fmul_rn(__fmul_rn(__fadd_rn(E[1], E[2]), 4.f), 1.0f / 64.0f);
    float v11 = __fmul_rn(__fmul_rn(__fadd_rn(O[1], O[2]), 4.f), 1.0f / 64.0f);
    const int64_t o0 = (int64_t)(2 * y) * a.dst_cols + 2 * x, o1 = o0 + a.dst_cols;
    if (a.minuend_ptrs || a.minuend) {
      v00 = pu_epilogue(a, b, o0, v00); v01 = pu_epilogue(a, b, o0 + 1, v01);
      v10 = pu_epilogue(a, b, o1, v10); v11 = pu_epilogue(a, b, o1 + 1, v11);
    }
    *reinterpret_cast<float2 *>(dst + o0) = make_float2(v00, v01);
    *reinterpret_cast<float2 *>(dst + o1) = make_float2(v10, v11);
    return;
  }
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int ox = 2 * x + (k & 1), oy = 2 * y + (k >> 1);
    if (ox >= a.dst_cols || oy >= a.dst_rows) continue;
    const int64_t o = (int64_t)oy * a.dst_cols + ox;
    dst[o] = pu_epilogue(a, b, o, pu_px(src, w, h, ox, oy));
  }
}

// ---- the small end of lpg's pyramid in one launch -------------------------------------------------
// lpg.cc:262-290 walks cv::pyrDown to a map of a few dozen pixels, scales it, raises it to the power p and walks cv::pyrUp
// back.  A level of a few thousand pixels is a few microseconds of launch latency and nothing else, so one 8-CTA cluster
// runs the levels from a small image down, the scalar steps, and back up to that size, with a cluster barrier between
// levels.  The levels are written inside this kernel by other CTAs: they are read with ld.global.cg (no L1).
// Per-output arithmetic = k_pyrdown's general path / pu_px.
__device__ __forceinline__ float pd_px_cg(const float *src, int w, int h, int dst_cols, int ox, int oy) {
  const int width0 = min((w - 3) / 2 + 1, dst_cols);
  const bool hsimd = ox >= 1 && ox < 1 + 4 * ((width0 - 1) / 4);
  const bool vsimd = ox < (dst_cols & ~3);
  int xx[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) xx[c] = border_idx(2 * ox - 2 + c, w, SSK_BORDER_REFLECT101);
  float r[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float *row = src + (int64_t)border_idx(2 * oy - 2 + k, h, SSK_BORDER_REFLECT101) * w;
    r[k] = pd_hform(__ldcg(row + xx[0]), __ldcg(row + xx[1]), __ldcg(row + xx[2]), __ldcg(row + xx[3]), __ldcg(row + xx[4]), hsimd);
  }
  const float a13 = __fadd_rn(r[1], r[3]);
  float v;
  if (vsimd) v = __fadd_rn(__fmul_rn(__fadd_rn(a13, r[2]), 4.f), __fadd_rn(__fadd_rn(r[0], r[4]), __fadd_rn(r[2], r[2])));
  else v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[2], 6.f), __fmul_rn(a13, 4.f)), r[0]), r[4]);
  return __fmul_rn(v, 1.0f / 256.0f);
}

// pu_row / pu_px with ld.global.cg
__device__ __forceinline__ float pu_row_cg(const float *p, int w, int ox) {
  const int x = min(ox >> 1, w - 1);
  const bool odd = (ox & 1) || (ox >> 1) >= w;
  if (odd) return x == w - 1 ? __fmul_rn(__ldcg(p + x), 8.f) : __fmul_rn(__fadd_rn(__ldcg(p + x), __ldcg(p + x + 1)), 4.f);
  if (x == 0) return __fadd_rn(__fmul_rn(__ldcg(p), 6.f), __fmul_rn(__ldcg(p + 1), 2.f));
  if (x == w - 1) return __fadd_rn(__ldcg(p + x - 1), __fmul_rn(__ldcg(p + x), 7.f));
  return __fadd_rn(__fadd_rn(__ldcg(p + x - 1), __fmul_rn(__ldcg(p + x), 6.f)), __ldcg(p + x + 1));
}
__device__ __forceinline__ float pu_px_cg(const float *src, int w, int h, int ox, int oy) {
  const int y = min(oy >> 1, h - 1);
  const bool odd = (oy & 1) || (oy >> 1) >= h;
  const int yd = min(y + 1, h - 1), yu = y == 0 ? 1 : y - 1;
  const float r1 = pu_row_cg(src + (int64_t)y * w, w, ox), r2 = pu_row_cg(src + (int64_t)yd * w, w, ox);
  float v;
  if (odd) v = __fmul_rn(__fadd_rn(r1, r2), 4.f);
  else v = __fadd_rn(__fadd_rn(__fmul_rn(r1, 6.f), pu_row_cg(src + (int64_t)yu * w, w, ox)), r2);
  return __fmul_rn(v, 1.0f / 64.0f);
}

constexpr int LT_CTAS = 8, LT_THREADS = 512;

__global__ void __launch_bounds__(LT_THREADS) k_lpg_tail(float *P, float *Q, int rows, int cols, int ndown, float scale, int ipow) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = (int)cluster.block_rank() * LT_THREADS + threadIdx.x, nthr = (int)cluster.num_blocks() * LT_THREADS;
  int hs[8], ws[8];
  hs[0] = rows; ws[0] = cols;
  for (int l = 0; l < ndown; ++l) { hs[l + 1] = (hs[l] + 1) / 2; ws[l + 1] = (ws[l] + 1) / 2; }
  float *cur = P, *oth = Q;
  for (int l = 0; l < ndown; ++l) {
    const int n = hs[l + 1] * ws[l + 1];
    const bool last = l == ndown - 1;
    for (int i = tid; i < n; i += nthr) {
      const int oy = i / ws[l + 1], ox = i - oy * ws[l + 1];
      float v = pd_px_cg(cur, ws[l], hs[l], ws[l + 1], ox, oy);
      if (last) {                                      // the smallest level: cv::multiply by (uscale - dscale), cv::pow (k_scale_ipow)
        v = __fmul_rn(v, scale);
        if (ipow > 1) {
          float a = 1.f, b = v;
          int p = ipow;
          while (p > 1) { if (p & 1) a = __fmul_rn(a, b); b = __fmul_rn(b, b); p >>= 1; }
          v = __fmul_rn(a, b);
        }
      }
      oth[i] = v;
    }
    cluster.sync();
    float *t = cur; cur = oth; oth = t;
  }
  for (int l = ndown - 1; l >= 0; --l) {
    const int n = hs[l] * ws[l];
    for (int i = tid; i < n; i += nthr) {
      const int oy = i / ws[l], ox = i - oy * ws[l];
      oth[i] = pu_px_cg(cur, ws[l + 1], hs[l + 1], ox, oy);
    }
    cluster.sync();
    float *t = cur; cur = oth; oth = t;
  }
}

// ---- W2: lpg --------------------------------------------------------------------------------------
// compute_lpg_5x5 (lpg.cc:60-129): alpha * laplacian^2 + beta * |gradient|^2 + eps with the 5x5 operators of the
// reference, single-rounded float operations in its evaluation order; the two border rows / columns repeat the
// nearest interior value.
__global__ void __launch_bounds__(256) k_lpg5x5(const float *__restrict__ src, int rows, int cols, float *__restrict__ dst,
                                                float alpha, float beta, float eps) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= cols || oy >= rows) return;
  const int x = min(max(ox, 2), cols - 3), y = min(max(oy, 2), rows - 3);
  auto r = [&](int dy, int dx) { return __ldg(src + (int64_t)(y + dy) * cols + (x + dx)); };
  auto add = [](float a, float b) { return __fadd_rn(a, b); };
  auto sub = [](float a, float b) { return __fsub_rn(a, b); };
  auto mul = [](float a, float b) { return __fmul_rn(a, b); };
  auto col5 = [&](int dx) { return add(add(add(add(r(-2, dx), mul(2.f, r(-1, dx))), mul(4.f, r(0, dx))), mul(2.f, r(1, dx))), r(2, dx)); };
  auto col3 = [&](int dx) { return add(add(r(-1, dx), mul(2.f, r(0, dx))), r(1, dx)); };
  auto row5 = [&](int dy) { return add(add(add(add(r(dy, -2), mul(2.f, r(dy, -1))), mul(4.f, r(dy, 0))), mul(2.f, r(dy, 1))), r(dy, 2)); };
  auto row3 = [&](int dy) { return add(add(r(dy, -1), mul(2.f, r(dy, 0))), r(dy, 1)); };
  const float gx = add(sub(col5(2), col5(-2)), mul(2.f, sub(col3(1), col3(-1))));
  const float gy = add(sub(row5(2), row5(-2)), mul(2.f, sub(row3(1), row3(-1))));
  const float grad = add(mul(gx, gx), mul(gy, gy));
  const float s4 = add(add(add(r(-1, 0), r(1, 0)), r(0, -1)), r(0, 1));
  const float d4 = add(add(add(r(-1, -1), r(-1, 1)), r(1, -1)), r(1, 1));
  const float f4 = add(add(add(r(-2, 0), r(2, 0)), r(0, -2)), r(0, 2));
  const float lap = sub(sub(sub(mul(16.f, r(0, 0)), mul(2.f, s4)), d4), f4);
  const float lapl = mul(lap, lap);
  dst[(int64_t)oy * cols + ox] = add(add(mul(alpha, lapl), mul(beta, grad)), eps);
}

// The same operator with its neighbours fused for the un-scaled case (dscale = 0): the channel average of a colour frame
// (reduce_color_channels, lpg.cc:246-248) is formed on load, the 36 x 20 source tile of a 32 x 16 output tile is staged once in
// shared memory (the plain kernel reads 25 + 12 global values per pixel) and the integer power (lpg.cc:270) is applied in
// registers: one read of the frame, one write of the map.  Every arithmetic step is the one of k_channel_avg / k_lpg5x5 /
// k_scale_ipow, in the same order.
constexpr int LF_W = 32, LF_H = 16;
template <int DEPTH>
__global__ void __launch_bounds__(256) k_lpg_fused(const Img im, float *__restrict__ dst, float alpha, float beta, float eps, int ipow) {
  __shared__ float s_t[LF_H + 4][LF_W + 4 + 1];
  const int bx = blockIdx.x * LF_W, by = blockIdx.y * LF_H;
  // origin of the staged tile = the first stencil centre of the tile minus 2.  Centres are clamped to [2, size - 3], so a last
  // tile narrower than 3 columns (or rows) has its centres left of (above) the tile: the origin follows them
  const int sx0 = min(max(bx, 2), im.cols - 3) - 2, sy0 = min(max(by, 2), im.rows - 3) - 2;
  const float inv_cn = (float)(1.0 / im.cn);
  for (int i = threadIdx.x; i < (LF_H + 4) * (LF_W + 4); i += 256) {
    const int r = i / (LF_W + 4), c = i - r * (LF_W + 4);
    const int gy = min(max(sy0 + r, 0), im.rows - 1), gx = min(max(sx0 + c, 0), im.cols - 1);
    float v = load_px<DEPTH>(im, gy, gx, 0);
    if (im.cn > 1) {
      for (int k = 1; k < im.cn; ++k) v = __fadd_rn(v, load_px<DEPTH>(im, gy, gx, k));
      v = __fmul_rn(v, inv_cn);
    }
    s_t[r][c] = v;
  }
  __syncthreads();
  auto add = [](float a, float b) { return __fadd_rn(a, b); };
  auto sub = [](float a, float b) { return __fsub_rn(a, b); };
  auto mul = [](float a, float b) { return __fmul_rn(a, b); };
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int ox = bx + (threadIdx.x & 31), oy = by + (threadIdx.x >> 5) + 8 * k;
    if (ox >= im.cols || oy >= im.rows) continue;
    const int x = min(max(ox, 2), im.cols - 3) - sx0, y = min(max(oy, 2), im.rows - 3) - sy0;   // tile coordinates of the stencil centre
    auto r = [&](int dy, int dx) { return s_t[y + dy][x + dx]; };
    auto col5 = [&](int dx) { return add(add(add(add(r(-2, dx), mul(2.f, r(-1, dx))), mul(4.f, r(0, dx))), mul(2.f, r(1, dx))), r(2, dx)); };
    auto col3 = [&](int dx) { return add(add(r(-1, dx), mul(2.f, r(0, dx))), r(1, dx)); };
    auto row5 = [&](int dy) { return add(add(add(add(r(dy, -2), mul(2.f, r(dy, -1))), mul(4.f, r(dy, 0))), mul(2.f, r(dy, 1))), r(dy, 2)); };
    auto row3 = [&](int dy) { return add(add(r(dy, -1), mul(2.f, r(dy, 0))), r(dy, 1)); };
    const float gx = add(sub(col5(2), col5(-2)), mul(2.f, sub(col3(1), col3(-1))));
    const float gy = add(sub(row5(2), row5(-2)), mul(2.f, sub(row3(1), row3(-1))));
    const float grad = add(mul(gx, gx), mul(gy, gy));
    const float s4 = add(add(add(r(-1, 0), r(1, 0)), r(0, -1)), r(0, 1));
    const float d4 = add(add(add(r(-1, -1), r(-1, 1)), r(1, -1)), r(1, 1));
    const float f4 = add(add(add(r(-2, 0), r(2, 0)), r(0, -2)), r(0, 2));
    const float lap = sub(sub(sub(mul(16.f, r(0, 0)), mul(2.f, s4)), d4), f4);
    float v = add(add(mul(alpha, mul(lap, lap)), mul(beta, grad)), eps);
    if (ipow > 1) {
      float a = 1.f, b = v;
      int p = ipow;
      while (p > 1) { if (p & 1) a = __fmul_rn(a, b); b = __fmul_rn(b, b); p >>= 1; }
      v = __fmul_rn(a, b);
    }
    dst[(int64_t)oy * im.cols + ox] = v;
  }
}

// v = cv::pow(v * scale, ipow) for integer ipow >= 1 (cv::multiply by a scalar, then iPow32f's square-and-multiply)
__global__ void __launch_bounds__(256) k_scale_ipow(float *buf, int64_t n, float scale, int apply_scale, int ipow) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float v = buf[i];
  if (apply_scale) v = __fmul_rn(v, scale);
  if (ipow > 1) {
    float a = 1.f, b = v;
    int p = ipow;
    while (p > 1) { if (p & 1) a = __fmul_rn(a, b); b = __fmul_rn(b, b); p >>= 1; }
    v = __fmul_rn(a, b);
  }
  buf[i] = v;
}

// ---- reference masks -----------------------------------------------------------------------------
// cv::pyrDown on CV_8UC1 (PyrDownInvoker with FixPtCast<uchar, 8>: integer taps, (sum + 128) >> 8, REFLECT101)
// followed by cv::compare(>= thresh): scaleImage's mask branch (c_frame_registration.cc:237-241)
__global__ void __launch_bounds__(256) k_pyrdown_mask_u8(const uint8_t *src, int64_t sstep, int rows, int cols, uint8_t *dst,
                                                         int drows, int dcols, int thresh) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= dcols || oy >= drows) return;
  const int kw[5] = {1, 4, 6, 4, 1};
  int sum = 0;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    const int yy = border_idx(2 * oy - 2 + dy, rows, SSK_BORDER_REFLECT101);
    int rsum = 0;
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) rsum += kw[dx] * src[(int64_t)yy * sstep + border_idx(2 * ox - 2 + dx, cols, SSK_BORDER_REFLECT101)];
    sum += kw[dy] * rsum;
  }
  const int v = (sum + 128) >> 8;
  dst[(int64_t)oy * dcols + ox] = v >= thresh ? 255 : 0;
}

// cv::resize(INTER_NEAREST) on CV_8UC1: sx = min(cvFloor(x * ifx), cols - 1), ifx = 1 / ((double)dcols / cols)
__global__ void __launch_bounds__(256) k_resize_nearest_u8(const uint8_t *src, int rows, int cols, uint8_t *dst, int drows, int dcols,
                                                           double ifx, double ify) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dcols || y >= drows) return;
  const int sx = min((int)floor(x * ifx), cols - 1), sy = min((int)floor(y * ify), rows - 1);
  dst[(int64_t)y * dcols + x] = src[(int64_t)sy * cols + sx];
}

// gradients are zeroed where the reference mask is zero (ecc_differentiate, ecc2.cc:162-166); *count += #nonzero
__global__ void __launch_bounds__(256) k_apply_refmask(const uint8_t *mask, int n, float *gx, float *gy, int *count) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  bool nz = false;
  if (i < n) {
    nz = mask[i] != 0;
    if (!nz && gx) { gx[i] = 0.f; gy[i] = 0.f; }
  }
  const unsigned b = __ballot_sync(0xffffffffu, nz);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

}  // namespace

int launch_pyrdown(const PyrDownArgs &a, cudaStream_t s) {
  SSK_REQUIRE(abs(a.dst_cols * 2 - a.src.cols) <= 2 && abs(a.dst_rows * 2 - a.src.rows) <= 2, "pyrDown: bad dstsize");
  dim3 grid(div_up(a.dst_cols, PD_OW), div_up(a.dst_rows, PD_OH), a.batch);
  if (a.src.depth == SSK_32F) k_pyrdown<SSK_32F><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_16U) k_pyrdown<SSK_16U><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_8U) k_pyrdown<SSK_8U><<<grid, 256, 0, s>>>(a);
  else { set_error("pyrDown: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_resize_area(const ResizeAreaArgs &a, cudaStream_t s) {
  SSK_REQUIRE(a.inv_scale_x > 0 && a.inv_scale_y > 0 && a.inv_scale_x <= 1.0 && a.inv_scale_y <= 1.0,
              "resize INTER_AREA: only down-scaling is implemented");
  SSK_REQUIRE(a.dst_cols >= 1 && a.dst_rows >= 1, "resize INTER_AREA: empty destination");
  const double scale_x = 1.0 / a.inv_scale_x, scale_y = 1.0 / a.inv_scale_y;   // cv::hal::resize
  int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);           // saturate_cast<int>(double)
  const bool fast = std::fabs(scale_x - iscale_x) < DBL_EPSILON && std::fabs(scale_y - iscale_y) < DBL_EPSILON;
  if (!fast) iscale_x = iscale_y = 0;
  dim3 grid(div_up(a.dst_cols, 32), div_up(a.dst_rows, 8), a.batch);
  // fractional scale on images large enough to matter: per-axis tables in stream-ordered scratch
  AreaAxis *tab = nullptr;
  if (!fast && (int64_t)a.dst_cols * a.dst_rows * a.batch >= 65536) {
    if (cudaMallocAsync(reinterpret_cast<void **>(&tab), sizeof(AreaAxis) * (size_t)(a.dst_cols + a.dst_rows), s) != cudaSuccess) { tab = nullptr; cudaGetLastError(); }
    else {
      k_area_axis_tab<<<div_up(a.dst_cols + a.dst_rows, 256), 256, 0, s>>>(a.dst_cols, a.dst_rows, scale_x, scale_y, a.src.cols, a.src.rows, tab);
      SSK_LAUNCH_CHECK();
    }
  }
  if (a.src.depth == SSK_32F) k_resize_area<SSK_32F><<<grid, 256, 0, s>>>(a, scale_x, scale_y, iscale_x, iscale_y, tab);
  else if (a.src.depth == SSK_16U) k_resize_area<SSK_16U><<<grid, 256, 0, s>>>(a, scale_x, scale_y, iscale_x, iscale_y, tab);
  else if (a.src.depth == SSK_8U) k_resize_area<SSK_8U><<<grid, 256, 0, s>>>(a, scale_x, scale_y, iscale_x, iscale_y, tab);
  else { if (tab) cudaFreeAsync(tab, s); set_error("resize INTER_AREA: unsupported depth"); return SSK_ERR_INVALID; }
  if (tab) cudaFreeAsync(tab, s);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_to_gray(const Img &src, const void *const *src_ptrs, float *dst, float *const *dst_ptrs, int batch, cudaStream_t s) {
  dim3 grid(div_up(src.cols, 32), div_up(src.rows, 8), batch);
  if (src.depth == SSK_32F) k_to_gray<SSK_32F><<<grid, 256, 0, s>>>(src, src_ptrs, dst, dst_ptrs);
  else if (src.depth == SSK_16U) k_to_gray<SSK_16U><<<grid, 256, 0, s>>>(src, src_ptrs, dst, dst_ptrs);
  else if (src.depth == SSK_8U) k_to_gray<SSK_8U><<<grid, 256, 0, s>>>(src, src_ptrs, dst, dst_ptrs);
  else { set_error("to_gray: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_sepfilter(const SepFilterArgs &a, cudaStream_t s) {
  SSK_REQUIRE((a.kxn & 1) && (a.kyn & 1) && a.kxn <= kMaxTaps && a.kyn <= kMaxTaps, "sepFilter2D: odd kernels up to 31 taps");
  const int cn = a.cn < 1 ? 1 : a.cn;
  SSK_REQUIRE(a.border == 0 || a.border == SSK_BORDER_REPLICATE || a.border == SSK_BORDER_REFLECT || a.border == SSK_BORDER_REFLECT101,
              "sepFilter2D: BORDER_REPLICATE, BORDER_REFLECT or BORDER_REFLECT101");
  dim3 grid(div_up(a.cols, SF_W), div_up(a.rows, SF_H), a.batch * cn);
  if (cn > 1 || (a.border && a.border != SSK_BORDER_REPLICATE)) k_sepfilter<0, 0><<<grid, 256, 0, s>>>(a);
  else if (a.kxn == 7 && a.kyn == 7) k_sepfilter<7, 7><<<grid, 256, 0, s>>>(a);     // Gaussian, sigma = 1
  else if (a.kxn == 9 && a.kyn == 9) k_sepfilter<9, 9><<<grid, 256, 0, s>>>(a);     // cv::GaussianBlur(sigma = 1) of CV_32F weights
  else if (a.kxn == 5 && a.kyn == 3) k_sepfilter<5, 3><<<grid, 256, 0, s>>>(a);     // ecc_differentiate, d/dx
  else if (a.kxn == 3 && a.kyn == 5) k_sepfilter<3, 5><<<grid, 256, 0, s>>>(a);     // ecc_differentiate, d/dy
  else k_sepfilter<0, 0><<<grid, 256, 0, s>>>(a);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

__global__ void __launch_bounds__(256) k_add_weighted(const float *__restrict__ src, double alpha, const float *__restrict__ lpass,
                                                      double beta, float *__restrict__ dst, int64_t n, int clamp, float outmin,
                                                      float outmax) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float v = __double2float_rn(__fma_rn((double)__ldg(src + i), alpha, __dmul_rn((double)__ldg(lpass + i), beta)));
    if (clamp) v = fmaxf(fminf(v, outmax), outmin);
    dst[i] = v;
  }
}

int launch_add_weighted(const float *src, double alpha, const float *lpass, double beta, float *dst, int64_t n, int clamp,
                        float outmin, float outmax, cudaStream_t s) {
  const int64_t blocks = (n + 255) / 256;
  k_add_weighted<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>(src, alpha, lpass, beta, dst, n, clamp, outmin, outmax);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int w1_num_blocks(int rows, int cols) { return div_up(cols, W1_W) * div_up(rows, W1_H); }

int launch_w1(const W1Args &a, cudaStream_t s) {
  SSK_REQUIRE(a.kradius >= 1 && a.kradius <= W1_RMAX, "local variance map: kradius 1..4");
  const int nblocks = w1_num_blocks(a.rows, a.cols);
  dim3 grid(div_up(a.cols, W1_W), div_up(a.rows, W1_H), a.batch);
  if (a.kradius == 1) k_w1_grad_r1<<<grid, 256, 0, s>>>(a, nblocks);
  else k_w1_grad<<<grid, 256, 0, s>>>(a, nblocks);
  SSK_LAUNCH_CHECK();
  k_w1_final<<<a.batch, 256, 0, s>>>(a, nblocks);
  SSK_LAUNCH_CHECK();
  if (a.out || a.out_ptrs) {
    W1Args u = a;     // geometry of the map that is up-sampled
    if (a.uscale > 0) {
      int ur, uc;
      w1_uscale_size(a.rows, a.cols, a.uscale, &ur, &uc);
      if (ur != a.rows || uc != a.cols) {
        SSK_REQUIRE(a.gmap2 || a.gmap2_ptrs, "local variance map: uscale scratch missing");
        ResizeAreaArgs ra = {};
        ra.src.data = a.gmap; ra.src.step = (int64_t)a.cols * 4; ra.src.rows = a.rows; ra.src.cols = a.cols;
        ra.src.depth = SSK_32F; ra.src.cn = 1; ra.src.scale = 1.f;
        ra.src_ptrs = reinterpret_cast<const void *const *>(a.gmap_ptrs);
        ra.dst = a.gmap2; ra.dst_ptrs = a.gmap2_ptrs; ra.dst_rows = ur; ra.dst_cols = uc; ra.batch = a.batch;
        ra.inv_scale_x = (double)uc / a.cols; ra.inv_scale_y = (double)ur / a.rows;   // cv::resize with an explicit dsize
        if (int e = launch_resize_area(ra, s)) return e;
        u.gmap = a.gmap2; u.gmap_ptrs = a.gmap2_ptrs; u.rows = ur; u.cols = uc;
      }
    }
    if (u.full_cols != u.cols || u.full_rows != u.rows) {
      SSK_REQUIRE(u.axis_tab, "local variance map: axis table scratch missing");
      if (!(u.axis_tab_built && *u.axis_tab_built)) {   // the tables depend on the geometry only: built once per handle
        const int xpad = (u.full_cols + 3) & ~3, ypad = (u.full_rows + 3) & ~3;
        int *xi = reinterpret_cast<int *>(u.axis_tab), *yi = xi + 2 * xpad;
        cudaMemsetAsync(u.axis_tab, 0, (size_t)(2 * xpad + 2 * ypad) * 4, s);
        k_w1_axis<<<div_up(u.full_cols, 256), 256, 0, s>>>(u.full_cols, u.cols, (double)u.cols / u.full_cols, xi, reinterpret_cast<float *>(xi + xpad));
        SSK_LAUNCH_CHECK();
        k_w1_axis<<<div_up(u.full_rows, 256), 256, 0, s>>>(u.full_rows, u.rows, (double)u.rows / u.full_rows, yi, reinterpret_cast<float *>(yi + ypad));
        SSK_LAUNCH_CHECK();
        if (u.axis_tab_built) *u.axis_tab_built = 1;
      }
    }
    if (u.full_cols == 2 * u.cols && u.full_rows == 2 * u.rows && (u.cols & 3) == 0 && u.cols >= 8 && u.rows >= 2 && !getenv("SSK_W1_UP_V4")) {
      dim3 g2(div_up(u.cols, 128), div_up(u.rows, 8), u.batch);
      k_w1_upsample2x_v8<<<g2, 256, 0, s>>>(u);
    } else if (u.full_cols == 2 * u.cols && u.full_rows == 2 * u.rows) {
      dim3 g2(div_up(u.full_cols, 128), div_up(u.full_rows, 16), u.batch);
      k_w1_upsample2x<<<g2, 256, 0, s>>>(u);
    } else {
      dim3 g2(div_up(u.full_cols, 128), div_up(u.full_rows, 8), u.batch);
      k_w1_upsample<<<g2, 256, 0, s>>>(u);
    }
    SSK_LAUNCH_CHECK();
  }
  return SSK_OK;
}

void w1_uscale_size(int rows, int cols, int uscale, int *urows, int *ucols) {
  // dscaleSize (c_local_variance_sharpness_measure.cc:15-25)
  for (int l = 0; l < uscale; ++l) {
    const int nc = (cols + 1) / 2, nr = (rows + 1) / 2;
    if (std::min(nc, nr) < 4) break;
    cols = nc; rows = nr;
  }
  *urows = rows; *ucols = cols;
}

int launch_lpg5x5(const float *src, int rows, int cols, float *dst, float alpha, float beta, float eps, cudaStream_t s) {
  SSK_REQUIRE(rows >= 5 && cols >= 5, "lpg: image smaller than 5x5 at the working scale");
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_lpg5x5<<<grid, 256, 0, s>>>(src, rows, cols, dst, alpha, beta, eps);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_lpg_fused(const Img &im, float *dst, float alpha, float beta, float eps, int ipow, cudaStream_t s) {
  SSK_REQUIRE(im.rows >= 5 && im.cols >= 5, "lpg: image smaller than 5x5 at the working scale");
  dim3 grid(div_up(im.cols, LF_W), div_up(im.rows, LF_H));
  if (im.depth == SSK_32F) k_lpg_fused<SSK_32F><<<grid, 256, 0, s>>>(im, dst, alpha, beta, eps, ipow);
  else if (im.depth == SSK_16U) k_lpg_fused<SSK_16U><<<grid, 256, 0, s>>>(im, dst, alpha, beta, eps, ipow);
  else if (im.depth == SSK_8U) k_lpg_fused<SSK_8U><<<grid, 256, 0, s>>>(im, dst, alpha, beta, eps, ipow);
  else { set_error("lpg: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_scale_ipow(float *buf, int64_t n, float scale, bool apply_scale, int ipow, cudaStream_t s) {
  k_scale_ipow<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(buf, n, scale, apply_scale ? 1 : 0, ipow);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

// P holds a rows x cols image; on return P holds pyrUp^ndown(pow(scale * pyrDown^ndown(P), ipow)) (ipow = 1: no power); Q is
// scratch of the same size.  One cluster of 8 CTAs.
int launch_lpg_tail(float *P, float *Q, int rows, int cols, int ndown, float scale, int ipow, cudaStream_t s) {
  SSK_REQUIRE(ndown >= 1 && ndown <= 7 && ipow >= 1, "lpg tail: bad geometry");
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(LT_CTAS);
  lc.blockDim = dim3(LT_THREADS);
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = LT_CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  void *args[7] = {(void *)&P, (void *)&Q, (void *)&rows, (void *)&cols, (void *)&ndown, (void *)&scale, (void *)&ipow};
  SSK_CUDA(cudaLaunchKernelExC(&lc, (const void *)k_lpg_tail, args));
  count_launch();
  return SSK_OK;
}

int launch_pyrup(const PyrUpArgs &a, cudaStream_t s) {
  SSK_REQUIRE(a.cols >= 2 && a.rows >= 2, "pyrUp: source smaller than 2x2");
  SSK_REQUIRE(abs(a.dst_cols - 2 * a.cols) <= 1 && abs(a.dst_rows - 2 * a.rows) <= 1, "pyrUp: bad dstsize");
  if (getenv("SSK_PYRUP_V1")) {       // A/B knob: one thread per output
    dim3 grid(div_up(a.dst_cols, 32), div_up(a.dst_rows, 8), a.batch);
    k_pyrup<<<grid, 256, 0, s>>>(a);
  } else {
    dim3 grid(div_up(div_up(a.dst_cols, 2), 32), div_up(div_up(a.dst_rows, 2), 8), a.batch);
    k_pyrup_2x2<<<grid, 256, 0, s>>>(a);
  }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_pyrdown_mask_u8(const uint8_t *src, int64_t sstep, int rows, int cols, uint8_t *dst, int drows, int dcols, int thresh,
                           cudaStream_t s) {
  dim3 grid(div_up(dcols, 32), div_up(drows, 8));
  k_pyrdown_mask_u8<<<grid, 256, 0, s>>>(src, sstep, rows, cols, dst, drows, dcols, thresh);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_resize_nearest_u8(const uint8_t *src, int rows, int cols, uint8_t *dst, int drows, int dcols, cudaStream_t s) {
  dim3 grid(div_up(dcols, 32), div_up(drows, 8));
  const double fx = (double)dcols / cols, fy = (double)drows / rows;
  k_resize_nearest_u8<<<grid, 256, 0, s>>>(src, rows, cols, dst, drows, dcols, 1.0 / fx, 1.0 / fy);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_apply_refmask(const uint8_t *mask, int n, float *gx, float *gy, int *d_count, cudaStream_t s) {
  k_apply_refmask<<<div_up(n, 256), 256, 0, s>>>(mask, n, gx, gy, d_count);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_erode5_u8(const uint8_t *src, int64_t sstep, uint8_t *dst, int64_t dstep, int rows, int cols,
                     int border_replicate, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_erode5_u8<<<grid, 256, 0, s>>>(src, sstep, dst, dstep, rows, cols, border_replicate);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk
