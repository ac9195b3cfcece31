// Device helpers shared by the fused warp+accumulate kernels (ssk_fused.cu: cp.async staging, interior + border-ring
// launch pair; ssk_fused_tma.cu: TMA staging, one launch, warps decoupled).  Included inside an anonymous namespace:
// every translation unit gets its own copy.
#pragma once
#include "ssk_warp.cuh"
#include <limits.h>

namespace ssk {

namespace {

constexpr int TW = 32, TH = 32;          // tile of the accumulator handled by one CTA of the staged kernel
constexpr int SWARPS = 4;                // warps per CTA, interior tiles (throughput: 8-row strips amortise the window fill)
#ifndef SSK_RING_WARPS
#define SSK_RING_WARPS 8
#endif
#ifndef SSK_RING_MINB
#define SSK_RING_MINB 2
#endif
constexpr int RING_WARPS = SSK_RING_WARPS;            // warps per CTA, border-ring tiles (heavier per frame than interior tiles: shorter chain per CTA)
constexpr int GSH = TH + 8;              // staged rows: tile + 3 taps + rounding + drift
constexpr int WWD = TW + 8;              // staged weight-tile row (floats)

__device__ __forceinline__ bool is_affine_like(int type) { return type != MAP_HOMOGRAPHY; }

template <int INTERP> struct Taps { static constexpr int N = INTERP == SSK_INTER_CUBIC ? 4 : INTERP == SSK_INTER_LINEAR ? 2 : 1;
                                    static constexpr int OFF = INTERP == SSK_INTER_CUBIC ? -1 : 0; };

// ------------------------------------------------------------------------------------------------
// generic per-pixel pieces (real function calls: rare paths must stay small)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ int border_idx_call(int p, int n, int border) { return border_idx(p, n, border); }

// cv::remap sample with run-time depth / interpolation / border, cv::remap's operation order (see ssk_common.cuh)
__device__ __noinline__ float sample_any(const Img &im, int c, float u, float v, int interp, int border, float bval,
                                         const float4 *cubic, const float *lanczos = nullptr) {
  if (interp == SSK_INTER_LANCZOS4) {
    if (im.depth == SSK_32F) return sample_lanczos4<SSK_32F>(im, c, u, v, border, bval, lanczos);
    if (im.depth == SSK_16U) return sample_lanczos4<SSK_16U>(im, c, u, v, border, bval, lanczos);
    return sample_lanczos4<SSK_8U>(im, c, u, v, border, bval, lanczos);
  }
  int ix, iy, fx = 0, fy = 0, n, off;
  float wx[4], wy[4];
  if (interp == SSK_INTER_NEAREST) {
    ix = __float2int_rn(u); iy = __float2int_rn(v); n = 1; off = 0; wx[0] = wy[0] = 1.f;
  } else {
    quant32(u, ix, fx);
    quant32(v, iy, fy);
    if (interp == SSK_INTER_LINEAR) {
      const float tx = (float)fx * 0.03125f, ty = (float)fy * 0.03125f;
      n = 2; off = 0; wx[0] = 1.0f - tx; wx[1] = tx; wy[0] = 1.0f - ty; wy[1] = ty;
    } else {
      const float4 cx = __ldg(cubic + fx), cy = __ldg(cubic + fy);
      n = 4; off = -1;
      wx[0] = cx.x; wx[1] = cx.y; wx[2] = cx.z; wx[3] = cx.w;
      wy[0] = cy.x; wy[1] = cy.y; wy[2] = cy.z; wy[3] = cy.w;
    }
  }
  float out = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < n; ++ky) {
    const int py = iy + off + ky;
    const int yy = (unsigned)py < (unsigned)im.rows ? py : border_idx_call(py, im.rows, border);
    float row = 0.f;
#pragma unroll 1
    for (int kx = 0; kx < n; ++kx) {
      const int px = ix + off + kx;
      const int xx = (unsigned)px < (unsigned)im.cols ? px : border_idx_call(px, im.cols, border);
      float s = bval;
      if (xx >= 0 && yy >= 0) {
        const char *p = static_cast<const char *>(im.data) + (int64_t)yy * im.step;
        if (im.depth == SSK_32F) s = __ldg(reinterpret_cast<const float *>(p) + xx * im.cn + c);
        else if (im.depth == SSK_16U) s = __fmul_rn((float)__ldg(reinterpret_cast<const uint16_t *>(p) + xx * im.cn + c), im.scale);
        else s = __fmul_rn((float)__ldg(reinterpret_cast<const uint8_t *>(p) + xx * im.cn + c), im.scale);
      }
      if (n == 1) return s;
      const float term = __fmul_rn(s, __fmul_rn(wy[ky], wx[kx]));
      if (n == 2) out = (ky == 0 && kx == 0) ? term : __fadd_rn(out, term);   // ((t00 + t01) + t10) + t11
      else row = __fadd_rn(row, term);
    }
    if (n == 4) out = __fadd_rn(out, row);
  }
  return out;
}

// mask(x, y) of base_remap: erode5x5(remap(all-255, interp, CONSTANT 0) >= 255) with border value 255
__device__ __noinline__ bool valid_eroded(const MapCoef &m, int interp, int x, int y, int cols, int rows, int src_cols,
                                          int src_rows, const Tables &tab) {
  float u, v;
  map_xy(m, (float)x, (float)y, u, v);
  if (is_affine_like(m.type)) {
    // neighbours within 2 px map within (2|a| + 2|b|) px of (u, v): if that stays inside the tap-safe interior every
    // one of the 25 pre-erosion flags is set
    float u1, v1, u2, v2;
    map_xy(m, (float)(x + 2), (float)y, u1, v1);
    map_xy(m, (float)x, (float)(y + 2), u2, v2);
    const float mg = interp == SSK_INTER_LANCZOS4 ? 5.f : 3.f;     // reach of the taps (Lanczos4: ix - 3 .. ix + 4)
    const float ru = fabsf(u1 - u) + fabsf(u2 - u) + mg, rv = fabsf(v1 - v) + fabsf(v2 - v) + mg;
    if (u - ru >= 0.f && v - rv >= 0.f && u + ru <= (float)(src_cols - 1) && v + rv <= (float)(src_rows - 1)) return true;
  }
#pragma unroll 1
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = y + dy;
    if ((unsigned)yy >= (unsigned)rows) continue;
#pragma unroll 1
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = x + dx;
      if ((unsigned)xx >= (unsigned)cols) continue;
      map_xy(m, (float)xx, (float)yy, u, v);
      if (!valid255_t(interp, u, v, src_cols, src_rows, tab)) return false;
    }
  }
  return true;
}

// Interior pixels of LINEAR / CUBIC warps (every tap inside the frame): the weight map and all channels of the frame are
// sampled with one set of tap coefficients and unrolled loads.  Operations and their order are sample_any's (the products
// wy * wx are formed once instead of once per channel and image).  N = taps per axis.
template <int N, int DEPTH>
__device__ __forceinline__ void interior_pixel(const WarpAccArgs &a, const FrameJob &job, const float (&cw)[N * N], int ix, int iy,
                                               bool weighted, float *A, float *W) {
  constexpr int OFF = N == 4 ? -1 : 0;
  typedef typename PixT<DEPTH>::type T;
  float wk = 1.f;
  if (weighted) {
    const float *wp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(job.weights) + (int64_t)(iy + OFF) * a.w_step) + (ix + OFF);
    float out = 0.f;
#pragma unroll
    for (int ky = 0; ky < N; ++ky) {
      const float *row = reinterpret_cast<const float *>(reinterpret_cast<const char *>(wp) + (int64_t)ky * a.w_step);
      float r = 0.f;
#pragma unroll
      for (int kx = 0; kx < N; ++kx) {
        const float term = __fmul_rn(__ldg(row + kx), cw[ky * N + kx]);
        if (N == 2) out = (ky == 0 && kx == 0) ? term : __fadd_rn(out, term);
        else r = __fadd_rn(r, term);
      }
      if (N == 4) out = __fadd_rn(out, r);
    }
    wk = out;
    if (!(wk > 0.f)) return;                      // c_frame_accumulation.cc:114
  }
  const float Wn = *W + wk;
  const float factor = weighted ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
  *W = Wn;
  const T *fp = reinterpret_cast<const T *>(static_cast<const char *>(job.frame) + (int64_t)(iy + OFF) * a.src_step) + (int64_t)(ix + OFF) * a.cn;
#pragma unroll 1
  for (int c = 0; c < a.cn; ++c) {
    float out = 0.f;
#pragma unroll
    for (int ky = 0; ky < N; ++ky) {
      const T *row = reinterpret_cast<const T *>(reinterpret_cast<const char *>(fp) + (int64_t)ky * a.src_step) + c;
      float r = 0.f;
#pragma unroll
      for (int kx = 0; kx < N; ++kx) {
        float sv;
        if (DEPTH == SSK_32F) sv = (float)__ldg(row + kx * a.cn);
        else sv = __fmul_rn((float)__ldg(row + kx * a.cn), a.scale);
        const float term = __fmul_rn(sv, cw[ky * N + kx]);
        if (N == 2) out = (ky == 0 && kx == 0) ? term : __fadd_rn(out, term);
        else r = __fadd_rn(r, term);
      }
      if (N == 4) out = __fadd_rn(out, r);
    }
    A[c] = fmaf(out - A[c], factor, A[c]);
  }
}

template <int N>
__device__ __noinline__ void interior_pixel_n(const WarpAccArgs &a, const Tables &tab, const FrameJob &job, int ix, int fx, int iy, int fy,
                                              bool weighted, float *A, float *W) {
  float wx[N], wy[N], cw[N * N];
  if (N == 2) {
    const float tx = (float)fx * 0.03125f, ty = (float)fy * 0.03125f;
    wx[0] = 1.0f - tx; wx[1] = tx; wy[0] = 1.0f - ty; wy[1] = ty;
  } else {
    const float4 cx = __ldg(tab.cubic + fx), cy = __ldg(tab.cubic + fy);
    wx[0] = cx.x; wx[1] = cx.y; wx[2 % N] = cx.z; wx[3 % N] = cx.w;
    wy[0] = cy.x; wy[1] = cy.y; wy[2 % N] = cy.z; wy[3 % N] = cy.w;
  }
#pragma unroll
  for (int ky = 0; ky < N; ++ky)
#pragma unroll
    for (int kx = 0; kx < N; ++kx) cw[ky * N + kx] = __fmul_rn(wy[ky], wx[kx]);
  if (a.depth == SSK_32F) interior_pixel<N, SSK_32F>(a, job, cw, ix, iy, weighted, A, W);
  else if (a.depth == SSK_16U) interior_pixel<N, SSK_16U>(a, job, cw, ix, iy, weighted, A, W);
  else interior_pixel<N, SSK_8U>(a, job, cw, ix, iy, weighted, A, W);
}

// one pixel of one frame through the generic path: A (cn values) and W are updated in place
__device__ __noinline__ void generic_pixel(const WarpAccArgs &a, const Tables &tab, const FrameJob &job, int x, int y,
                                           float *A, float *W) {
  const MapCoef m = job.map;
  if (!valid_eroded(m, a.interp, x, y, a.cols, a.rows, a.src_cols, a.src_rows, tab)) return;
  float u, v;
  map_xy(m, (float)x, (float)y, u, v);
  Img im;
  im.rows = a.src_rows; im.cols = a.src_cols;
  const bool weighted = a.use_weights && job.weights != nullptr;
  float wk = 1.f;
  if (weighted) {
    im.data = job.weights; im.step = a.w_step; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
    wk = sample_any(im, 0, u, v, a.interp, SSK_BORDER_CONSTANT, 0.f, tab.cubic, tab.lanczos);
    if (!(wk > 0.f)) return;                      // c_frame_accumulation.cc:114
  }
  im.data = job.frame; im.step = a.src_step; im.depth = a.depth; im.cn = a.cn; im.scale = a.scale;
  const float Wn = *W + wk;
  const float factor = weighted ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
  *W = Wn;
  for (int c = 0; c < a.cn; ++c) {
    const float I = sample_any(im, c, u, v, a.interp, a.border, a.bval[c], tab.cubic, tab.lanczos);
    A[c] = fmaf(I - A[c], factor, A[c]);
  }
}

// generic_pixel with the interior fast path in front: used by k_fused_generic only (multi-channel frames, projective maps).
// The staged kernels keep calling generic_pixel itself for their rare fall-backs: their code and its layout stay as tuned
// (routing them through this function cost the TMA kernel 5 %, measured).
__device__ __forceinline__ void generic_pixel_fast(const WarpAccArgs &a, const Tables &tab, const FrameJob &job, int x, int y,
                                                   float *A, float *W) {
  if (a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC) {
    const MapCoef m = job.map;
    float u, v;
    map_xy(m, (float)x, (float)y, u, v);
    int ix, fx, iy, fy;
    quant32(u, ix, fx);
    quant32(v, iy, fy);
    const int off = a.interp == SSK_INTER_CUBIC ? -1 : 0, n = a.interp == SSK_INTER_CUBIC ? 4 : 2;
    if (ix + off >= 0 && iy + off >= 0 && ix + off + n <= a.src_cols && iy + off + n <= a.src_rows) {
      if (!valid_eroded(m, a.interp, x, y, a.cols, a.rows, a.src_cols, a.src_rows, tab)) return;
      const bool wtd = a.use_weights && job.weights != nullptr;
      if (n == 4) interior_pixel_n<4>(a, tab, job, ix, fx, iy, fy, wtd, A, W);
      else interior_pixel_n<2>(a, tab, job, ix, fx, iy, fy, wtd, A, W);
      return;
    }
  }
  generic_pixel(a, tab, job, x, y, A, W);
}

// ------------------------------------------------------------------------------------------------
// k_fused_generic: one thread per pixel of a list of tiles (the border ring, or the whole image)
// ------------------------------------------------------------------------------------------------
struct TileList {            // tiles of TW x TH pixels; `ring` selects the border ring of an ntx x nty tiling
  int ntx, nty, ring;
};

__device__ __forceinline__ void tile_of_block(const TileList &t, int b, int &tx, int &ty) {
  if (!t.ring) { tx = b % t.ntx; ty = b / t.ntx; return; }
  if (b < t.ntx) { tx = b; ty = 0; return; }
  b -= t.ntx;
  if (b < t.ntx) { tx = b; ty = t.nty - 1; return; }
  b -= t.ntx;
  if (b < t.nty - 2) { tx = 0; ty = 1 + b; return; }
  b -= t.nty - 2;
  tx = t.ntx - 1; ty = 1 + b;
}

// ------------------------------------------------------------------------------------------------
// k_fused_staged
// ------------------------------------------------------------------------------------------------
template <int DEPTH> struct StageGeom {
  static constexpr int ES = (DEPTH == SSK_32F ? 4 : DEPTH == SSK_16U ? 2 : 1);
  static constexpr int ALIGN = 16 / ES;                                        // elements per 16-byte chunk
  static constexpr int WD = ((TW + 4 + ALIGN + ALIGN - 1) / ALIGN) * ALIGN;    // staged frame row (elements)
  static constexpr int ROWB = WD * ES;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NKEEP> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NKEEP)); }

// ---- TMA staging (interior tiles, 32F frame + weight map): one elected thread asks the tensor copy engine for the
// whole GSH x WD window of both images; completion is counted on an mbarrier every thread of the CTA waits on.
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SSK_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SSK_MBAR_DONE;\n"
      "bra SSK_MBAR_WAIT;\n"
      "SSK_MBAR_DONE:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned smem_dst, const void *tmap, int x, int y, unsigned mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
               "l"(tmap), "r"(x), "r"(y), "r"(mbar)
               : "memory");
}
// the tensor maps live in global memory and are rewritten by the host between launches: make the copy engine re-read them
__device__ __forceinline__ void tmap_acquire(const void *tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}

__device__ __forceinline__ void tmap_prefetch(const void *tmap) { asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory"); }

struct StagePlan { int staged, sx0, sy0, sxw; };   // staged: tile safe and footprint fits; sxw: weight-tile origin
// Plans of all frames of a launch are computed once per CTA (one frame per thread) and kept in shared memory.
constexpr int KPLAN = 256;                          // frames per launch (longer batches are split by the launcher)
struct PackedPlan { short sx0, sy0, sxw, staged; }; // staged: 1 staged, 0 generic path, -1 frame dropped by registration


template <int DEPTH>
__device__ __forceinline__ float lds_px(const unsigned char *p, float scale) {
  if (DEPTH == SSK_32F) return *reinterpret_cast<const float *>(p);
  if (DEPTH == SSK_16U) return __fmul_rn((float)*reinterpret_cast<const uint16_t *>(p), scale);
  return __fmul_rn((float)*p, scale);
}

// Column-invariant part of an affine-like map: the per-row evaluation keeps the reference's operand order
// (bit-identical to map_xy) without a type switch in the hot loop.
template <int MT> struct ColMap {
  float a, b, c, d, e, f, g, h;
  __device__ __forceinline__ ColMap(const MapCoef &m, float x) {
    a = b = c = d = e = f = g = h = 0.f;
    if (MT == MAP_TRANSLATION) { a = __fadd_rn(x, m.c[0]); b = m.c[1]; }
    else if (MT == MAP_AFFINE) { a = __fmul_rn(m.c[0], x); b = m.c[1]; c = m.c[2]; d = __fmul_rn(m.c[3], x); e = m.c[4]; f = m.c[5]; }
    else {  // euclidean: xx = x - Cx
      const float xx = __fsub_rn(x, m.c[5]);
      a = __fmul_rn(m.c[1], xx); b = __fmul_rn(m.c[2], xx); c = m.c[0]; d = m.c[1]; e = m.c[2]; f = m.c[3]; g = m.c[4]; h = m.c[6];
    }
  }
  __device__ __forceinline__ void operator()(float y, float &u, float &v) const {
    if (MT == MAP_TRANSLATION) { u = a; v = __fadd_rn(y, b); }
    else if (MT == MAP_AFFINE) {
      u = __fadd_rn(__fadd_rn(a, __fmul_rn(b, y)), c);
      v = __fadd_rn(__fadd_rn(d, __fmul_rn(e, y)), f);
    } else {
      const float yy = __fsub_rn(y, h);
      u = __fadd_rn(__fmul_rn(c, __fsub_rn(a, __fmul_rn(e, yy))), f);
      v = __fadd_rn(__fmul_rn(c, __fadd_rn(b, __fmul_rn(d, yy))), g);
    }
  }
};

template <int INTERP> struct RollS {
  static constexpr int N = Taps<INTERP>::N;
  float f[N][N], w[N][N];    // frame / weight windows, rows in rotating slots
  int ix, iy;                // source anchor of the windows
  const unsigned char *pf;   // staged frame row that enters the window next
  const float *pw;           // staged weight row that enters the window next
};

// One output pixel from the staged tiles: slide the windows one row down (or re-anchor them), interpolate the weight
// and the frame, update the running weighted mean held in shared memory (predicated, straight-line).
// Bicubic is evaluated separably with FMAs (row sums first): it differs from cv::remap's 16-product sum in the last
// ulp only, far inside the 1e-4 stack tolerance; bilinear keeps cv::remap's exact order.
template <int DEPTH, int INTERP, bool WEIGHTS, int MT, int J>
__device__ __forceinline__ void roll_pixel_s(RollS<INTERP> &R, const ColMap<MT> &cm, float y, const unsigned char *s_f,
                                             const float *s_g, float scale, const StagePlan &pl, const float4 *s_cubic,
                                             float *s_acc_px, float *s_w_px, bool ok = true) {
  typedef StageGeom<DEPTH> G;
  constexpr int N = Taps<INTERP>::N, OFF = Taps<INTERP>::OFF;
  float u, v;
  cm(y, u, v);
  int ix, iy, fx = 0, fy = 0;
  if (INTERP == SSK_INTER_NEAREST) { ix = __float2int_rn(u); iy = __float2int_rn(v); }
  else { quant32(u, ix, fx); quant32(v, iy, fy); }
  if (ix != R.ix || iy != R.iy + 1) {
    R.pf = s_f + (iy + OFF - pl.sy0) * G::ROWB + (ix + OFF - pl.sx0) * G::ES;
    R.pw = s_g + (iy + OFF - pl.sy0) * WWD + (ix + OFF - pl.sxw);
#pragma unroll
    for (int r = 0; r < N - 1; ++r) {
#pragma unroll
      for (int q = 0; q < N; ++q) {
        R.f[(J + r) % N][q] = lds_px<DEPTH>(R.pf + q * G::ES, scale);
        if (WEIGHTS) R.w[(J + r) % N][q] = R.pw[q];
      }
      R.pf += G::ROWB; R.pw += WWD;
    }
  }
#pragma unroll
  for (int q = 0; q < N; ++q) {
    R.f[(J + N - 1) % N][q] = lds_px<DEPTH>(R.pf + q * G::ES, scale);
    if (WEIGHTS) R.w[(J + N - 1) % N][q] = R.pw[q];
  }
  R.pf += G::ROWB; R.pw += WWD;
  R.ix = ix; R.iy = iy;

  float I, wk = 1.f;
  if (INTERP == SSK_INTER_CUBIC) {
    const float4 cx = s_cubic[fx], cy = s_cubic[fy];
    float rs[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float *t = R.f[(J + r) % N];
      rs[r] = fmaf(t[3], cx.w, fmaf(t[2], cx.z, fmaf(t[1], cx.y, t[0] * cx.x)));
    }
    I = fmaf(rs[3], cy.w, fmaf(rs[2], cy.z, fmaf(rs[1], cy.y, rs[0] * cy.x)));
    if (WEIGHTS) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float *t = R.w[(J + r) % N];
        rs[r] = fmaf(t[3], cx.w, fmaf(t[2], cx.z, fmaf(t[1], cx.y, t[0] * cx.x)));
      }
      wk = fmaf(rs[3], cy.w, fmaf(rs[2], cy.z, fmaf(rs[1], cy.y, rs[0] * cy.x)));
    }
  } else if (INTERP == SSK_INTER_LINEAR) {
    // cv::remapBilinear's exact order: ((S00*w00 + S01*w01) + S10*w10) + S11*w11
    const float tx = (float)fx * 0.03125f, ty = (float)fy * 0.03125f;
    const float w00 = __fmul_rn(1.0f - ty, 1.0f - tx), w01 = __fmul_rn(1.0f - ty, tx), w10 = __fmul_rn(ty, 1.0f - tx), w11 = __fmul_rn(ty, tx);
    const float *f0 = R.f[J % N], *f1 = R.f[(J + 1) % N];
    I = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(f0[0], w00), __fmul_rn(f0[1], w01)), __fmul_rn(f1[0], w10)), __fmul_rn(f1[1], w11));
    if (WEIGHTS) {
      const float *g0 = R.w[J % N], *g1 = R.w[(J + 1) % N];
      wk = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g0[0], w00), __fmul_rn(g0[1], w01)), __fmul_rn(g1[0], w10)), __fmul_rn(g1[1], w11));
    }
  } else {
    I = R.f[0][0];
    if (WEIGHTS) wk = R.w[0][0];
  }
  const float W0 = *s_w_px, A = *s_acc_px;
  const float Wn = W0 + wk;
  const float factor = WEIGHTS ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
  const bool upd = ok && (!WEIGHTS || wk > 0.f);  // eroded validity mask (ring tiles); c_frame_accumulation.cc:114
  *s_w_px = upd ? Wn : W0;
  *s_acc_px = upd ? fmaf(I - A, factor, A) : A;
}

// ---- packed (frame, weight) pairs: one FFMA2 / FMUL2 interpolates both images (sm_100 f32x2 arithmetic) ----
typedef unsigned long long pair_t;
__device__ __forceinline__ pair_t pk2(float lo, float hi) { pair_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ pair_t mul2s(pair_t a, float s) {
  pair_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(pk2(s, s))); return r;
}
__device__ __forceinline__ pair_t fma2s(pair_t a, float s, pair_t c) {
  pair_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(pk2(s, s)), "l"(c)); return r;
}
__device__ __forceinline__ float lds_f32(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float4 lds_f32x4(unsigned addr) {
  float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

struct RollC2 {
  pair_t fw[4][4];           // (frame, weight) taps, rows in rotating slots
  int ix, iy;                // source anchor of the window
  const unsigned char *pf;   // staged frame row that enters the window next
  const float *pw;           // staged weight row that enters the window next
};

// Bicubic + weight map specialisation of roll_pixel_s: the frame and weight windows travel as packed pairs, so the
// 20 multiply-adds of the separable bicubic serve both images; the accumulator tile and the coefficient table are
// addressed through 32-bit shared addresses computed once per thread; the running-mean factor w/(W+w) uses the
// approximate reciprocal (the reference itself is built with -ffast-math; <= 2 ulp on the factor).
template <int DEPTH, int MT, int J>
__device__ __forceinline__ void roll_pixel_c2(RollC2 &R, const ColMap<MT> &cm, float y, const unsigned char *s_f, const float *s_g,
                                              float scale, const StagePlan &pl, unsigned cub_a, unsigned acc_a, unsigned w_a, bool ok = true) {
  typedef StageGeom<DEPTH> G;
  float u, v;
  cm(y, u, v);
  const int su = __float2int_rn(__fmul_rn(u, 32.0f)), sv = __float2int_rn(__fmul_rn(v, 32.0f));
  const int ix = su >> kInterBits, iy = sv >> kInterBits;
  if (ix != R.ix || iy != R.iy + 1) {
    R.pf = s_f + (iy - 1 - pl.sy0) * G::ROWB + (ix - 1 - pl.sx0) * G::ES;
    R.pw = s_g + (iy - 1 - pl.sy0) * WWD + (ix - 1 - pl.sxw);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int q = 0; q < 4; ++q) R.fw[(J + r) % 4][q] = pk2(lds_px<DEPTH>(R.pf + q * G::ES, scale), R.pw[q]);
      R.pf += G::ROWB; R.pw += WWD;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) R.fw[(J + 3) % 4][q] = pk2(lds_px<DEPTH>(R.pf + q * G::ES, scale), R.pw[q]);
  R.pf += G::ROWB; R.pw += WWD;
  R.ix = ix; R.iy = iy;

  const float4 cx = lds_f32x4(cub_a + ((su & (kInterTab - 1)) << 4)), cy = lds_f32x4(cub_a + ((sv & (kInterTab - 1)) << 4));
  pair_t rs[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const pair_t *t = R.fw[(J + r) % 4];
    rs[r] = fma2s(t[3], cx.w, fma2s(t[2], cx.z, fma2s(t[1], cx.y, mul2s(t[0], cx.x))));
  }
  const pair_t res = fma2s(rs[3], cy.w, fma2s(rs[2], cy.z, fma2s(rs[1], cy.y, mul2s(rs[0], cy.x))));
  float I, wk;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(I), "=f"(wk) : "l"(res));
  const float W0 = lds_f32(w_a), A = lds_f32(acc_a);
  const float Wn = W0 + wk;
  float rW;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rW) : "f"(Wn));
  // wk <= Wn, so the exact factor is <= 1: the clamp keeps a denormal weight sum (flushed: rcp = inf) from poisoning
  // the pixel with inf / NaN; elsewhere the approximate reciprocal is within 2 ulp of wk / Wn
  const float factor = fminf(wk * rW, 1.0f);
  if (ok && wk > 0.f) {              // eroded validity mask (ring tiles); c_frame_accumulation.cc:114
    sts_f32(w_a, Wn);
    sts_f32(acc_a, fmaf(I - A, factor, A));
  }
}

// cv::borderInterpolate for the modes whose mapped index stays near the border (a single reflection suffices for
// the few pixels of overhang a tile can have); -1: BORDER_CONSTANT (use the border value)
__device__ __forceinline__ int bmap(int p, int n, int border) {
  if ((unsigned)p < (unsigned)n) return p;
  if (border == SSK_BORDER_REPLICATE) return p < 0 ? 0 : n - 1;
  if (border == SSK_BORDER_REFLECT101) return p < 0 ? -p : 2 * (n - 1) - p;
  if (border == SSK_BORDER_REFLECT) return p < 0 ? -p - 1 : 2 * n - 1 - p;
  return -1;
}

}  // namespace

}  // namespace ssk
