// K5: fused sub-pixel warp + validity mask + weighted running-mean accumulation, plus the un-fused
// building blocks (cv::remap equivalent, mask remap + erode, accumulator add / compute, Bayer average).
//
// Reference semantics reproduced here:
//   c_frame_registration::base_remap        core/proc/image_registration/c_frame_registration.cc:1265-1386
//   multiply_weights                        core/pipeline/c_image_stacking_pipeline/c_image_stacking_pipeline.cc:108-138,1704-1714
//   _weighted_average_update                core/average/c_frame_accumulation.cc:20-129
//   _bayer_accumulate / compute             core/average/c_frame_accumulation.cc:988-1126, 1205-1238
// (the fused warp + accumulate kernels live in ssk_fused.cu)
#include "ssk_warp.cuh"

namespace ssk {

// ------------------------------------------------------------------------------------------------
// un-fused cv::remap (CV_32F, 1..4 channels)
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void job_coords(const MapCoef &m, const float2 *rmap, int64_t rmap_step, int x, int y,
                                           float &u, float &v) {
  if (rmap) {
    const float2 p = *reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(rmap) + (int64_t)y * rmap_step + (int64_t)x * 8);
    u = p.x; v = p.y;
  } else {
    map_xy(m, (float)x, (float)y, u, v);
  }
}

__global__ void __launch_bounds__(256) k_remap(const RemapArgs a, const Tables tab) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.cols || y >= a.rows) return;
  float u, v;
  job_coords(a.map, a.rmap, a.rmap_step, x, y, u, v);
  float *d = reinterpret_cast<float *>(reinterpret_cast<char *>(a.dst) + (int64_t)y * a.dst_step) + x * a.src.cn;
  if (a.border == SSK_BORDER_TRANSPARENT) {
    // dst is left untouched when the anchor tap is outside the source (cv::remap behaviour, cvmodel.py)
    int ix, iy, f;
    if (a.interp == SSK_INTER_NEAREST) { ix = __float2int_rn(u); iy = __float2int_rn(v); }
    else { quant32(u, ix, f); quant32(v, iy, f); }
    if ((unsigned)ix >= (unsigned)a.src.cols || (unsigned)iy >= (unsigned)a.src.rows) return;
  }
  for (int c = 0; c < a.src.cn; ++c) {
    float r;
    if (a.interp == SSK_INTER_CUBIC) r = sample_cubic<SSK_32F>(a.src, c, u, v, a.border, a.bval[c], tab.cubic);
    else if (a.interp == SSK_INTER_LANCZOS4) r = sample_lanczos4<SSK_32F>(a.src, c, u, v, a.border, a.bval[c], tab.lanczos);
    else if (a.interp == SSK_INTER_NEAREST) r = sample_nearest<SSK_32F>(a.src, c, u, v, a.border, a.bval[c]);
    else r = sample_linear<SSK_32F>(a.src, c, u, v, a.border, a.bval[c]);
    d[c] = r;
  }
}

// pre-erode validity of a remapped 8U mask (or of an all-255 source)
__global__ void __launch_bounds__(256) k_mask_pre(const RemapMaskArgs a, const Tables tab) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.cols || y >= a.rows) return;
  float u, v;
  job_coords(a.map, a.rmap, a.rmap_step, x, y, u, v);
  bool ok;
  if (!a.src_mask) {
    ok = valid255_t(a.interp, u, v, a.src_cols, a.src_rows, tab);
  } else if (a.interp == SSK_INTER_NEAREST) {
    const int ix = __float2int_rn(u), iy = __float2int_rn(v);
    ok = (unsigned)ix < (unsigned)a.src_cols && (unsigned)iy < (unsigned)a.src_rows &&
         a.src_mask[(int64_t)iy * a.src_mask_step + ix] >= 255;
  } else {
    // fixed-point remap of an arbitrary 8U mask, BORDER_CONSTANT 0
    int ix, fx, iy, fy;
    quant32(u, ix, fx);
    quant32(v, iy, fy);
    int S = 0;
    if (a.interp == SSK_INTER_LANCZOS4) {
      const short *w = tab.lanczos_itab + ((fy << kInterBits) + fx) * 64;
      for (int ky = 0; ky < 8; ++ky)
        for (int kx = 0; kx < 8; ++kx) {
          const int xx = ix - 3 + kx, yy = iy - 3 + ky;
          if ((unsigned)xx < (unsigned)a.src_cols && (unsigned)yy < (unsigned)a.src_rows)
            S += w[ky * 8 + kx] * (int)a.src_mask[(int64_t)yy * a.src_mask_step + xx];
        }
    } else if (a.interp == SSK_INTER_CUBIC) {
      const short *w = tab.cubic_itab + ((fy << kInterBits) + fx) * 16;
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx) {
          const int xx = ix - 1 + kx, yy = iy - 1 + ky;
          if ((unsigned)xx < (unsigned)a.src_cols && (unsigned)yy < (unsigned)a.src_rows)
            S += w[ky * 4 + kx] * (int)a.src_mask[(int64_t)yy * a.src_mask_step + xx];
        }
    } else {
      const int wx[2] = {32 - fx, fx}, wyv[2] = {32 - fy, fy};
      for (int ky = 0; ky < 2; ++ky)
        for (int kx = 0; kx < 2; ++kx) {
          const int xx = ix + kx, yy = iy + ky;
          if ((unsigned)xx < (unsigned)a.src_cols && (unsigned)yy < (unsigned)a.src_rows)
            S += wx[kx] * wyv[ky] * 32 * (int)a.src_mask[(int64_t)yy * a.src_mask_step + xx];
        }
    }
    const int val = (S + (1 << (kCoefBits - 1))) >> kCoefBits;
    ok = val >= 255;
  }
  a.tmp[(int64_t)y * a.cols + x] = ok ? 255 : 0;
}

__global__ void __launch_bounds__(256) k_erode5(const uint8_t *src, int rows, int cols, uint8_t *dst, int64_t dst_step) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  uint8_t m = 255;
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = y + dy;
    if ((unsigned)yy >= (unsigned)rows) continue;   // border value 255
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = x + dx;
      if ((unsigned)xx >= (unsigned)cols) continue;
      m = min(m, src[(int64_t)yy * cols + xx]);
    }
  }
  dst[(int64_t)y * dst_step + x] = m;
}

}  // namespace

int launch_remap(const RemapArgs &a, const Tables &tab, cudaStream_t s) {
  SSK_REQUIRE(a.src.depth == SSK_32F && a.src.cn >= 1 && a.src.cn <= 4, "remap: CV_32F source with 1..4 channels");
  SSK_REQUIRE(a.interp == SSK_INTER_NEAREST || a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC || a.interp == SSK_INTER_LANCZOS4,
              "remap: interpolation must be NEAREST, LINEAR, CUBIC, AREA or LANCZOS4");
  dim3 grid(div_up(a.cols, 32), div_up(a.rows, 8));
  k_remap<<<grid, 256, 0, s>>>(a, tab);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_remap_mask(const RemapMaskArgs &a, const Tables &tab, cudaStream_t s) {
  SSK_REQUIRE(a.interp == SSK_INTER_NEAREST || a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC || a.interp == SSK_INTER_LANCZOS4,
              "remap: interpolation must be NEAREST, LINEAR, CUBIC, AREA or LANCZOS4");
  dim3 grid(div_up(a.cols, 32), div_up(a.rows, 8));
  k_mask_pre<<<grid, 256, 0, s>>>(a, tab);
  SSK_LAUNCH_CHECK();
  k_erode5<<<grid, 256, 0, s>>>(a.tmp, a.rows, a.cols, a.dst, a.dst_step);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

// ------------------------------------------------------------------------------------------------
// D1: ellipsoid derotation map.  Double precision, single-rounded operations in the reference's evaluation order
// (no FMA contraction), so that the hit / visibility decisions at the limb match the CPU code.
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds(double a, double b) { return __dsub_rn(a, b); }

__global__ void __launch_bounds__(256) k_ellipsoid_remap(const EllipsoidArgs a) {
  const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ix >= a.cols || iy >= a.rows) return;
  const int64_t o = (int64_t)iy * a.cols + ix;
  float2 m = make_float2((float)ix, (float)iy);
  uint8_t mask = 0;
  float w = 0.f;
  if (iy >= a.by && iy <= a.by + a.bh && ix >= a.bx && ix <= a.bx + a.bw) {
    // ellipsoid_from_cart2d (ellipsoid.h:107-158) under the target pose R2
    const double *R = a.R2;
    const double xs = ds((double)ix, a.cx), ys = ds((double)iy, a.cy);
    const double x_stat = da(dm(R[0], xs), dm(R[3], ys));
    const double y_stat = da(dm(R[1], xs), dm(R[4], ys));
    const double z_stat = da(dm(R[2], xs), dm(R[5], ys));
    const double rzx = R[6], rzy = R[7], rzz = R[8];
    const double iA = __ddiv_rn(1.0, dm(a.A, a.A)), iB = __ddiv_rn(1.0, dm(a.B, a.B)), iC = __ddiv_rn(1.0, dm(a.C, a.C));
    const double K2 = da(da(dm(dm(rzx, rzx), iA), dm(dm(rzy, rzy), iB)), dm(dm(rzz, rzz), iC));
    const double K1 = da(da(dm(dm(x_stat, rzx), iA), dm(dm(y_stat, rzy), iB)), dm(dm(z_stat, rzz), iC));
    const double K0 = ds(da(da(dm(dm(x_stat, x_stat), iA), dm(dm(y_stat, y_stat), iB)), dm(dm(z_stat, z_stat), iC)), 1.0);
    const double disc = ds(dm(K1, K1), dm(K2, K0));
    if (!(disc < 0.0)) {
      mask = 255;
      const double sq = __dsqrt_rn(disc);
      const double zs1 = __ddiv_rn(ds(-K1, sq), K2), zs2 = __ddiv_rn(da(-K1, sq), K2);
      const double zs = fmin(zs1, zs2);
      // v = R2^T * (xs, ys, zs); vcam = R1 * v  (cv::Matx products: sums in index order)
      const double vx = da(da(dm(R[0], xs), dm(R[3], ys)), dm(R[6], zs));
      const double vy = da(da(dm(R[1], xs), dm(R[4], ys)), dm(R[7], zs));
      const double vz = da(da(dm(R[2], xs), dm(R[5], ys)), dm(R[8], zs));
      const double *Q = a.R1;
      const double px = da(da(dm(Q[0], vx), dm(Q[1], vy)), dm(Q[2], vz));
      const double py = da(da(dm(Q[3], vx), dm(Q[4], vy)), dm(Q[5], vz));
      const double pz = da(da(dm(Q[6], vx), dm(Q[7], vy)), dm(Q[8], vz));
      if (pz <= 0.0) m = make_float2((float)da(px, a.cx), (float)da(py, a.cy));
      else m = make_float2(-1.f, -1.f);
      // limb weight (ellipsoid.cc:250-272), rows / columns of the crop box proper
      if (iy < a.by + a.bh && ix < a.bx + a.bw) {
        const double dx = xs, dy = ys;
        const double xx = dm(da(dm(dx, a.ca), dm(dy, a.sa)), __ddiv_rn(1.0, a.A));
        const double yy = dm(da(dm(-dx, a.sa), dm(dy, a.ca)), __ddiv_rn(1.0, a.B));
        const double rr = da(dm(xx, xx), dm(yy, yy));
        if (rr <= 1.0) w = (float)dm(a.wscale, __dsqrt_rn(fmax(0.0, ds(1.0, rr))));
      }
    }
  }
  a.rmap[o] = m;
  a.rmask[o] = mask;
  a.wmap[o] = w;
}

}  // namespace

namespace {
__global__ void __launch_bounds__(256) k_jdr_weights(float *w, const float *lpg_map, const uint8_t *rmask, const uint8_t *mask,
                                                     int64_t mask_step, int rows, int cols, int is_master) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int64_t o = (int64_t)y * cols + x;
  float v = w[o];
  if (v < 1e-5f) v = 0.f;                      // current_weights.setTo(0, current_weights < 1e-5): scalar taken as float
  if (lpg_map) v = __fmul_rn(v, lpg_map[o]);   // cv::multiply(current_weights, lpg_map)
  if (is_master && !rmask[o]) v = 1.f;         // setTo(1, ~rmask)
  if (mask && !mask[(int64_t)y * mask_step + x]) v = 0.f;
  w[o] = v;
}
}  // namespace

namespace {
// cv::remap(src, dst, rmap, INTER_LINEAR, BORDER_TRANSPARENT) of a dense CV_32FC1 image applied in place, one sample: the
// destination keeps its own value where the anchor tap of the map lies outside the source (k_remap's rule)
__device__ __forceinline__ float remap_transparent_inplace(const Img &im, float2 m, int x, int y) {
  int ix, iy, f;
  quant32(m.x, ix, f);
  quant32(m.y, iy, f);
  if ((unsigned)ix >= (unsigned)im.cols || (unsigned)iy >= (unsigned)im.rows) return load_px<SSK_32F>(im, y, x, 0);
  return sample_linear<SSK_32F>(im, 0, m.x, m.y, SSK_BORDER_TRANSPARENT, 0.f);
}

// c_jdr_pipeline.cc:1207-1227 in one pass: w = remap(wmap, rmap, LINEAR, CONSTANT 0); w < 1e-5 -> 0; w *= remap(lpg, rmap,
// LINEAR, TRANSPARENT in place); master frame: 1 outside the disk; 0 under the frame mask.  Same per-sample arithmetic as the
// k_remap / k_jdr_weights chain it replaces.
__global__ void __launch_bounds__(256) k_jdr_weights_fused(const float *wpre, const float *lpg_map, const float2 *rmap, const uint8_t *rmask,
                                                           const uint8_t *mask, int64_t mask_step, int rows, int cols, int is_master, float *w) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int64_t o = (int64_t)y * cols + x;
  const float2 m = rmap[o];
  Img im;
  im.data = wpre; im.step = (int64_t)cols * 4; im.rows = rows; im.cols = cols; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
  float v = sample_linear<SSK_32F>(im, 0, m.x, m.y, SSK_BORDER_CONSTANT, 0.f);
  if (v < 1e-5f) v = 0.f;
  if (lpg_map) {
    im.data = lpg_map;
    v = __fmul_rn(v, remap_transparent_inplace(im, m, x, y));
  }
  if (is_master && !rmask[o]) v = 1.f;
  if (mask && !mask[(int64_t)y * mask_step + x]) v = 0.f;
  w[o] = v;
}

// derotation of the frame (cv::remap LINEAR / TRANSPARENT in place, c_jdr_pipeline.cc:1231-1234) fused with
// c_weigthed_average::add(frame, weights) (c_frame_accumulation.cc:20-129, CV_32FC1 weights)
__global__ void __launch_bounds__(256) k_jdr_remap_add(const float *frame, const float2 *rmap, const float *weights, int rows, int cols,
                                                       float *acc, float *wacc) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int64_t o = (int64_t)y * cols + x;
  const float wnew = weights[o];
  if (!(wnew > 0.f)) return;
  Img im;
  im.data = frame; im.step = (int64_t)cols * 4; im.rows = rows; im.cols = cols; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
  const float I = remap_transparent_inplace(im, rmap[o], x, y);
  const float Wn = wacc[o] + wnew;
  const float factor = __fdiv_rn(wnew, Wn);
  wacc[o] = Wn;
  const float A = acc[o];
  acc[o] = fmaf(I - A, factor, A);
}
}  // namespace

int launch_jdr_weights_fused(const float *wpre, const float *lpg_map, const float2 *rmap, const uint8_t *rmask, const uint8_t *mask,
                             int64_t mask_step, int rows, int cols, int is_master, float *w, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_jdr_weights_fused<<<grid, 256, 0, s>>>(wpre, lpg_map, rmap, rmask, mask, mask_step, rows, cols, is_master, w);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_jdr_remap_add(const float *frame, const float2 *rmap, const float *weights, int rows, int cols, float *acc, float *wacc,
                         cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_jdr_remap_add<<<grid, 256, 0, s>>>(frame, rmap, weights, rows, cols, acc, wacc);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_jdr_weights(float *w, const float *lpg_map, const uint8_t *rmask, const uint8_t *mask, int64_t mask_step, int rows,
                       int cols, int is_master, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_jdr_weights<<<grid, 256, 0, s>>>(w, lpg_map, rmask, mask, mask_step, rows, cols, is_master);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_ellipsoid_remap(const EllipsoidArgs &a, cudaStream_t s) {
  dim3 grid(div_up(a.cols, 32), div_up(a.rows, 8));
  k_ellipsoid_remap<<<grid, 256, 0, s>>>(a);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

// ------------------------------------------------------------------------------------------------
// accumulator kernels
// ------------------------------------------------------------------------------------------------
namespace {

template <int DEPTH>
__global__ void __launch_bounds__(256) k_acc_add(const AccAddArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.src.cols || y >= a.src.rows) return;
  float wnew = 1.f;
  if (a.wtype == SSK_8UC1) {
    if (!static_cast<const uint8_t *>(a.weights)[(int64_t)y * a.w_step + x]) return;
  } else if (a.wtype == SSK_32FC1) {
    wnew = *reinterpret_cast<const float *>(static_cast<const char *>(a.weights) + (int64_t)y * a.w_step + (int64_t)x * 4);
    if (!(wnew > 0.f)) return;
  }
  const int64_t p = (int64_t)y * a.src.cols + x;
  const float Wn = a.wacc[p] + wnew;
  const float factor = a.wtype == SSK_32FC1 ? __fdiv_rn(wnew, Wn) : __fdiv_rn(1.0f, Wn);
  a.wacc[p] = Wn;
  for (int c = 0; c < a.src.cn; ++c) {
    const float I = load_px<DEPTH>(a.src, y, x, c);
    const float A = a.acc[p * a.src.cn + c];
    a.acc[p * a.src.cn + c] = fmaf(I - A, factor, A);
  }
}

__global__ void __launch_bounds__(256) k_acc_compute(const float *acc, const float *wacc, int rows, int cols, int cn,
                                                     float dscale, float *avg, int64_t avg_step, uint8_t *mask,
                                                     int64_t mask_step) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int64_t p = (int64_t)y * cols + x;
  if (avg) {
    float *d = reinterpret_cast<float *>(reinterpret_cast<char *>(avg) + (int64_t)y * avg_step) + x * cn;
    for (int c = 0; c < cn; ++c) d[c] = dscale == 1.f ? acc[p * cn + c] : acc[p * cn + c] * dscale;
  }
  if (mask) mask[(int64_t)y * mask_step + x] = wacc[p] > 0.f ? 255 : 0;
}

__global__ void __launch_bounds__(256) k_acc_sum_form(float *acc, const float *wacc, int64_t npix, int cn, int to_sum) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float w = wacc[p];
  for (int c = 0; c < cn; ++c) {
    const float a = acc[p * cn + c];
    acc[p * cn + c] = to_sum ? a * w : (w > 0.f ? __fdiv_rn(a, w) : 0.f);
  }
}

__device__ __forceinline__ int bayer_channel(int colorid, int y, int x) {
  // c_frame_accumulation.cc:1262-1334; B=0 G=1 R=2 (c_frame_accumulation.h:230-234)
  const int q = ((y & 1) << 1) | (x & 1);
  switch (colorid) {
    case SSK_COLORID_BAYER_RGGB: return q == 0 ? 2 : q == 3 ? 0 : 1;
    case SSK_COLORID_BAYER_GRBG: return q == 1 ? 2 : q == 2 ? 0 : 1;
    case SSK_COLORID_BAYER_GBRG: return q == 2 ? 2 : q == 1 ? 0 : 1;
    default: /* BGGR */          return q == 3 ? 2 : q == 0 ? 0 : 1;
  }
}

template <int DEPTH>
__global__ void __launch_bounds__(256) k_bayer_add(const BayerAccArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int rows = a.src.rows, cols = a.src.cols;
  if (x >= cols || y >= rows) return;
  double w = 1.0;
  if (a.wtype == SSK_8UC1) {
    if (!static_cast<const uint8_t *>(a.weights)[(int64_t)y * a.w_step + x]) return;
  } else if (a.wtype == SSK_32FC1) {
    w = *reinterpret_cast<const float *>(static_cast<const char *>(a.weights) + (int64_t)y * a.w_step + (int64_t)x * 4);
  }
  const int64_t p = ((int64_t)y * cols + x) * 3;
  float acc[3] = {a.acc[p], a.acc[p + 1], a.acc[p + 2]};
  float cn[3] = {a.cntr[p], a.cntr[p + 1], a.cntr[p + 2]};
  // the pattern table only covers the even-sized part of the image (c_frame_accumulation.cc:1262-1334)
  const int prow = rows & ~1, pcol = cols & ~1;
  if (!a.have_map) {
    const int cc = (y < prow && x < pcol) ? bayer_channel(a.colorid, y, x) : 0;
    const float s = load_px<DEPTH>(a.src, y, x, 0);
    if (a.wtype == SSK_32FC1) { acc[cc] += s * (float)w; cn[cc] += (float)w; }
    else { acc[cc] += s; cn[cc] += 1.f; }
  } else {
    float u, v;
    job_coords(a.map, a.rmap, a.rmap_step, x, y, u, v);
    const int sx = (int)u, sy = (int)v;   // truncation toward zero, as the reference's (int) cast
    if (!(sx >= 0 && sx < cols - 1 && sy >= 0 && sy < rows - 1)) return;
    const double ax = (double)((float)(sx + 1) - u), ay = (double)((float)(sy + 1) - v);
    const double bx = (double)(u - (float)sx), by = (double)(v - (float)sy);
    const double sw[4] = {ax * ay * w, bx * ay * w, ax * by * w, bx * by * w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = sy + (k >> 1), xx = sx + (k & 1);
      const int cc = (yy < prow && xx < pcol) ? bayer_channel(a.colorid, yy, xx) : 0;
      const double s = (double)load_px<DEPTH>(a.src, yy, xx, 0);
      acc[cc] = (float)((double)acc[cc] + s * sw[k]);
      cn[cc] = (float)((double)cn[cc] + sw[k]);
    }
  }
  a.acc[p] = acc[0]; a.acc[p + 1] = acc[1]; a.acc[p + 2] = acc[2];
  a.cntr[p] = cn[0]; a.cntr[p + 1] = cn[1]; a.cntr[p + 2] = cn[2];
}

__global__ void __launch_bounds__(256) k_bayer_compute(const float *acc, const float *cntr, int rows, int cols, float *avg,
                                                       int64_t avg_step, uint8_t *mask, int64_t mask_step) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int64_t p = ((int64_t)y * cols + x) * 3;
  bool any = false;
  float *d = avg ? reinterpret_cast<float *>(reinterpret_cast<char *>(avg) + (int64_t)y * avg_step) + x * 3 : nullptr;
  for (int c = 0; c < 3; ++c) {
    const float n = cntr[p + c];
    any = any || n > 0.f;
    if (d) d[c] = n > 0.f ? __fdiv_rn(acc[p + c], n) : 0.f;
  }
  if (mask) mask[(int64_t)y * mask_step + x] = any ? 255 : 0;
}

}  // namespace

int launch_acc_add(const AccAddArgs &a, cudaStream_t s) {
  dim3 grid(div_up(a.src.cols, 32), div_up(a.src.rows, 8));
  if (a.src.depth == SSK_32F) k_acc_add<SSK_32F><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_16U) k_acc_add<SSK_16U><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_8U) k_acc_add<SSK_8U><<<grid, 256, 0, s>>>(a);
  else { set_error("acc_add: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_acc_compute(const float *acc, const float *wacc, int rows, int cols, int cn, float dscale, float *avg,
                       int64_t avg_step, uint8_t *mask, int64_t mask_step, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_acc_compute<<<grid, 256, 0, s>>>(acc, wacc, rows, cols, cn, dscale, avg, avg_step, mask, mask_step);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_acc_sum_form(float *acc, const float *wacc, int64_t npix, int cn, int to_sum, cudaStream_t s) {
  k_acc_sum_form<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(acc, wacc, npix, cn, to_sum);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_bayer_add(const BayerAccArgs &a, cudaStream_t s) {
  dim3 grid(div_up(a.src.cols, 32), div_up(a.src.rows, 8));
  if (a.src.depth == SSK_32F) k_bayer_add<SSK_32F><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_16U) k_bayer_add<SSK_16U><<<grid, 256, 0, s>>>(a);
  else if (a.src.depth == SSK_8U) k_bayer_add<SSK_8U><<<grid, 256, 0, s>>>(a);
  else { set_error("bayer_add: unsupported depth"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_bayer_compute(const float *acc, const float *cntr, int rows, int cols, float *avg, int64_t avg_step,
                         uint8_t *mask, int64_t mask_step, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_bayer_compute<<<grid, 256, 0, s>>>(acc, cntr, rows, cols, avg, avg_step, mask, mask_step);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk
