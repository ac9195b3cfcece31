// K5: fused sub-pixel warp + validity mask + weighted running-mean accumulation, plus the un-fused
// building blocks (cv::remap equivalent, mask remap + erode, accumulator add / compute, Bayer average).
//
// Reference semantics reproduced here:
//   c_frame_registration::base_remap        core/proc/image_registration/c_frame_registration.cc:1265-1386
//   multiply_weights                        core/pipeline/c_image_stacking_pipeline/c_image_stacking_pipeline.cc:108-138,1704-1714
//   _weighted_average_update                core/average/c_frame_accumulation.cc:20-129
//   _bayer_accumulate / compute             core/average/c_frame_accumulation.cc:988-1126, 1205-1238
// One pass per frame: the frame (and its weight map) are read once, the accumulators are read and written
// once per *batch* (they stay in registers across the frames of a batch).
#include "ssk_warp.cuh"

namespace ssk {

namespace {

constexpr int TW = 128;   // tile width  (32 lanes x 4 px, 16-byte accesses)
constexpr int TH = 8;     // tile height (8 warps)
constexpr int HALO = 2;   // 5x5 erosion
constexpr int FW = TW + 2 * HALO;
constexpr int FH = TH + 2 * HALO;

template <int DEPTH, int INTERP>
__device__ __forceinline__ float sample(const Img &im, int c, float u, float v, int border, float bval,
                                        const Tables &tab) {
  if (INTERP == SSK_INTER_CUBIC) return sample_cubic<DEPTH>(im, c, u, v, border, bval, tab.cubic);
  if (INTERP == SSK_INTER_NEAREST) return sample_nearest<DEPTH>(im, c, u, v, border, bval);
  return sample_linear<DEPTH>(im, c, u, v, border, bval);
}

__device__ __forceinline__ bool is_affine_like(int type) { return type != MAP_HOMOGRAPHY; }

template <int DEPTH, int CN, int INTERP, bool WEIGHTS>
__global__ void __launch_bounds__(256) k_warp_acc(const WarpAccArgs a, const Tables tab) {
  __shared__ uint8_t s_valid[FH][FW + 4];

  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int bx0 = blockIdx.x * TW, by0 = blockIdx.y * TH;
  const int x0 = bx0 + lane * 4, y = by0 + wy;
  const bool row_ok = y < a.rows;
  const bool vec = (a.cols & 3) == 0;
  const int npx = !row_ok ? 0 : min(4, a.cols - x0);   // <= 0 when the thread is outside

  float A[4][CN], W[4];
  const int64_t pix0 = (int64_t)y * a.cols + x0;
  if (npx > 0) {
    if (vec) {
      const float4 w4 = *reinterpret_cast<const float4 *>(a.wacc + pix0);
      W[0] = w4.x; W[1] = w4.y; W[2] = w4.z; W[3] = w4.w;
      float tmp[4 * CN];
#pragma unroll
      for (int k = 0; k < CN; ++k)
        *reinterpret_cast<float4 *>(tmp + 4 * k) = *reinterpret_cast<const float4 *>(a.acc + pix0 * CN + 4 * k);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) A[i][c] = tmp[i * CN + c];
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        W[i] = i < npx ? a.wacc[pix0 + i] : 0.f;
#pragma unroll
        for (int c = 0; c < CN; ++c) A[i][c] = i < npx ? a.acc[(pix0 + i) * CN + c] : 0.f;
      }
    }
  }

  Img src;
  src.step = a.src_step; src.rows = a.src_rows; src.cols = a.src_cols; src.depth = DEPTH; src.cn = CN; src.scale = a.scale;
  Img wim;
  wim.step = a.w_step; wim.rows = a.src_rows; wim.cols = a.src_cols; wim.depth = SSK_32F; wim.cn = 1; wim.scale = 1.f;

  // tile + halo rectangle clipped to the image: its corners decide whether the whole tile maps into the
  // safe interior of the frame (then every pixel is valid and the 5x5 erosion cannot remove any).
  const int cx0 = max(bx0 - HALO, 0), cy0 = max(by0 - HALO, 0);
  const int cx1 = min(bx0 + TW - 1 + HALO, a.cols - 1), cy1 = min(by0 + TH - 1 + HALO, a.rows - 1);

  for (int j = 0; j < a.njobs; ++j) {
    const FrameJob &job = a.jobs[j];
    if (!job.ok) continue;                       // block-uniform
    const MapCoef m = job.map;
    src.data = job.frame;
    wim.data = job.weights;

    bool need_flags = true;
    if (is_affine_like(m.type)) {
      float u, v;
      bool safe = true;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        map_xy(m, (float)((k & 1) ? cx1 : cx0), (float)((k & 2) ? cy1 : cy0), u, v);
        safe = safe && u >= 3.f && v >= 3.f && u <= (float)(a.src_cols - 4) && v <= (float)(a.src_rows - 4);
      }
      need_flags = !safe;
    }

    if (need_flags) {
      for (int k = threadIdx.x; k < FH * FW; k += blockDim.x) {
        const int fy = k / FW, fx = k - fy * FW;
        const int gx = bx0 - HALO + fx, gy = by0 - HALO + fy;
        uint8_t ok = 1;                          // outside the image: erode border value 255
        if (gx >= 0 && gy >= 0 && gx < a.cols && gy < a.rows) {
          float u, v;
          map_xy(m, (float)gx, (float)gy, u, v);
          ok = valid255(INTERP, u, v, a.src_cols, a.src_rows, tab.cubic_itab) ? 1 : 0;
        }
        s_valid[fy][fx] = ok;
      }
      __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= npx) continue;
      if (need_flags) {
        bool mv = true;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) mv = mv && s_valid[wy + dy][lane * 4 + i + dx];
        if (!mv) continue;
      }
      float u, v;
      map_xy(m, (float)(x0 + i), (float)y, u, v);
      float wnew = 1.f;
      if (WEIGHTS) {
        wnew = sample<SSK_32F, INTERP>(wim, 0, u, v, SSK_BORDER_CONSTANT, 0.f, tab);
        if (!(wnew > 0.f)) continue;             // c_frame_accumulation.cc:114
      }
      const float Wn = W[i] + wnew;
      const float factor = WEIGHTS ? __fdiv_rn(wnew, Wn) : __fdiv_rn(1.0f, Wn);
      W[i] = Wn;
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const float I = sample<DEPTH, INTERP>(src, c, u, v, a.border, a.bval[c], tab);
        A[i][c] = fmaf(I - A[i][c], factor, A[i][c]);
      }
    }
    if (need_flags) __syncthreads();
  }

  if (npx > 0) {
    if (vec) {
      *reinterpret_cast<float4 *>(a.wacc + pix0) = make_float4(W[0], W[1], W[2], W[3]);
      float tmp[4 * CN];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) tmp[i * CN + c] = A[i][c];
#pragma unroll
      for (int k = 0; k < CN; ++k)
        *reinterpret_cast<float4 *>(a.acc + pix0 * CN + 4 * k) = *reinterpret_cast<const float4 *>(tmp + 4 * k);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= npx) continue;
        a.wacc[pix0 + i] = W[i];
#pragma unroll
        for (int c = 0; c < CN; ++c) a.acc[(pix0 + i) * CN + c] = A[i][c];
      }
    }
  }
}

template <int DEPTH, int CN>
int launch_wa_dc(const WarpAccArgs &a, const Tables &tab, cudaStream_t s) {
  dim3 grid(div_up(a.cols, TW), div_up(a.rows, TH)), block(256);
#define SSK_WA(I, WGT) k_warp_acc<DEPTH, CN, I, WGT><<<grid, block, 0, s>>>(a, tab)
  if (a.use_weights) {
    if (a.interp == SSK_INTER_CUBIC) SSK_WA(SSK_INTER_CUBIC, true);
    else if (a.interp == SSK_INTER_NEAREST) SSK_WA(SSK_INTER_NEAREST, true);
    else SSK_WA(SSK_INTER_LINEAR, true);
  } else {
    if (a.interp == SSK_INTER_CUBIC) SSK_WA(SSK_INTER_CUBIC, false);
    else if (a.interp == SSK_INTER_NEAREST) SSK_WA(SSK_INTER_NEAREST, false);
    else SSK_WA(SSK_INTER_LINEAR, false);
  }
#undef SSK_WA
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace

int launch_warp_accumulate(const WarpAccArgs &a, const Tables &tab, cudaStream_t s) {
  SSK_REQUIRE(a.interp == SSK_INTER_NEAREST || a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC,
              "warp_accumulate: interpolation must be NEAREST, LINEAR or CUBIC");
  SSK_REQUIRE(a.border != SSK_BORDER_TRANSPARENT, "warp_accumulate: BORDER_TRANSPARENT is not meaningful here");
  SSK_REQUIRE(a.cn == 1 || a.cn == 3, "warp_accumulate: 1 or 3 channels");
  if (a.cn == 1) {
    if (a.depth == SSK_32F) return launch_wa_dc<SSK_32F, 1>(a, tab, s);
    if (a.depth == SSK_16U) return launch_wa_dc<SSK_16U, 1>(a, tab, s);
    if (a.depth == SSK_8U) return launch_wa_dc<SSK_8U, 1>(a, tab, s);
  } else {
    if (a.depth == SSK_32F) return launch_wa_dc<SSK_32F, 3>(a, tab, s);
    if (a.depth == SSK_16U) return launch_wa_dc<SSK_16U, 3>(a, tab, s);
    if (a.depth == SSK_8U) return launch_wa_dc<SSK_8U, 3>(a, tab, s);
  }
  set_error("warp_accumulate: unsupported frame depth");
  return SSK_ERR_INVALID;
}

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
    ok = valid255(a.interp, u, v, a.src_cols, a.src_rows, tab.cubic_itab);
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
    if (a.interp == SSK_INTER_CUBIC) {
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
  dim3 grid(div_up(a.cols, 32), div_up(a.rows, 8));
  k_remap<<<grid, 256, 0, s>>>(a, tab);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_remap_mask(const RemapMaskArgs &a, const Tables &tab, cudaStream_t s) {
  dim3 grid(div_up(a.cols, 32), div_up(a.rows, 8));
  k_mask_pre<<<grid, 256, 0, s>>>(a, tab);
  SSK_LAUNCH_CHECK();
  k_erode5<<<grid, 256, 0, s>>>(a.tmp, a.rows, a.cols, a.dst, a.dst_step);
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
