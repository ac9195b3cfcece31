// K5: fused sub-pixel warp + validity mask + weight-map warp + weighted running-mean accumulation for a batch of
// frames (one launch per batch; the accumulators are read and written once per batch, every frame and weight map
// is read once).
//
// Reference semantics reproduced here:
//   c_frame_registration::base_remap   core/proc/image_registration/c_frame_registration.cc:1265-1386
//        frame: cv::remap(interp, border); mask: remap(all-255, interp, CONSTANT 0) >= 255, erode 5x5 (border 255)
//   weights: custom_remap(weights, interp, BORDER_CONSTANT), weights *= mask/255
//                                      core/pipeline/c_image_stacking_pipeline/c_image_stacking_pipeline.cc:1653-1660, 1704-1714
//   _weighted_average_update           core/average/c_frame_accumulation.cc:20-129
//
// Two kernels share the work (separate kernels keep each instruction footprint inside the 32 KB L1.5 I-cache):
//   k_fused_staged   interior tiles, single channel: the tile's source footprint of frame j+1 is copied to shared
//                    memory with cp.async (LDGSTS.128) while frame j is interpolated from shared memory with a rolling
//                    register window (4 new taps per pixel instead of 16); the running mean / weight of the tile stay
//                    in shared memory for the whole batch and move to / from HBM with 16-byte accesses.
//   k_fused_generic  the ring of tiles along the image border (per-tap cv::borderInterpolate, eroded validity mask),
//                    multi-channel frames, projective maps: one thread per pixel, accumulators in registers.
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
                                         const float4 *cubic) {
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
                                          int src_rows, const short *itab) {
  float u, v;
  map_xy(m, (float)x, (float)y, u, v);
  if (is_affine_like(m.type)) {
    // neighbours within 2 px map within (2|a| + 2|b|) px of (u, v): if that stays inside the tap-safe interior every
    // one of the 25 pre-erosion flags is set
    float u1, v1, u2, v2;
    map_xy(m, (float)(x + 2), (float)y, u1, v1);
    map_xy(m, (float)x, (float)(y + 2), u2, v2);
    const float ru = fabsf(u1 - u) + fabsf(u2 - u) + 3.f, rv = fabsf(v1 - v) + fabsf(v2 - v) + 3.f;
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
      if (!valid255(interp, u, v, src_cols, src_rows, itab)) return false;
    }
  }
  return true;
}

// one pixel of one frame through the generic path: A (cn values) and W are updated in place
__device__ __noinline__ void generic_pixel(const WarpAccArgs &a, const Tables &tab, const FrameJob &job, int x, int y,
                                           float *A, float *W) {
  const MapCoef m = job.map;
  if (!valid_eroded(m, a.interp, x, y, a.cols, a.rows, a.src_cols, a.src_rows, tab.cubic_itab)) return;
  float u, v;
  map_xy(m, (float)x, (float)y, u, v);
  Img im;
  im.rows = a.src_rows; im.cols = a.src_cols;
  const bool weighted = a.use_weights && job.weights != nullptr;
  float wk = 1.f;
  if (weighted) {
    im.data = job.weights; im.step = a.w_step; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
    wk = sample_any(im, 0, u, v, a.interp, SSK_BORDER_CONSTANT, 0.f, tab.cubic);
    if (!(wk > 0.f)) return;                      // c_frame_accumulation.cc:114
  }
  im.data = job.frame; im.step = a.src_step; im.depth = a.depth; im.cn = a.cn; im.scale = a.scale;
  const float Wn = *W + wk;
  const float factor = weighted ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
  *W = Wn;
  for (int c = 0; c < a.cn; ++c) {
    const float I = sample_any(im, c, u, v, a.interp, a.border, a.bval[c], tab.cubic);
    A[c] = fmaf(I - A[c], factor, A[c]);
  }
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

__global__ void __launch_bounds__(256) k_fused_generic(const __grid_constant__ WarpAccArgs a, const __grid_constant__ Tables tab, const TileList tl) {
  int tx, ty;
  tile_of_block(tl, blockIdx.x >> 2, tx, ty);                 // 4 CTAs of 32 x 8 pixels per tile
  const int x = tx * TW + (threadIdx.x & 31);
  const int y = ty * TH + (blockIdx.x & 3) * 8 + (threadIdx.x >> 5);
  if (x >= a.cols || y >= a.rows) return;
  const int64_t p = (int64_t)y * a.cols + x;
  float A[4], W = a.wacc[p];
  for (int c = 0; c < a.cn; ++c) A[c] = a.acc[p * a.cn + c];
#pragma unroll 1
  for (int j = 0; j < a.njobs; ++j) {
    if (!a.jobs[j].ok) continue;
    generic_pixel(a, tab, a.jobs[j], x, y, A, &W);
  }
  a.wacc[p] = W;
  for (int c = 0; c < a.cn; ++c) a.acc[p * a.cn + c] = A[c];
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

struct StagePlan { int staged, sx0, sy0, sxw; };   // staged: tile safe and footprint fits; sxw: weight-tile origin
// Plans of all frames of a launch are computed once per CTA (one frame per thread) and kept in shared memory.
constexpr int KPLAN = 256;                          // frames per launch (longer batches are split by the launcher)
struct PackedPlan { short sx0, sy0, sxw, staged; }; // staged: 1 staged, 0 generic path, -1 frame dropped by registration

// Footprint of the tile in the frame (block-uniform).  es / align describe the frame element type.
__device__ __noinline__ StagePlan plan_stage(const MapCoef &m, int bx0, int by0, const WarpAccArgs &a, int align, int wd) {
  StagePlan p; p.staged = 0; p.sx0 = p.sy0 = p.sxw = 0;
  if (!a.stage_aligned || !is_affine_like(m.type)) return p;
  const int cx0 = max(bx0 - 2, 0), cy0 = max(by0 - 2, 0);
  const int cx1 = min(bx0 + TW + 1, a.cols - 1), cy1 = min(by0 + TH + 1, a.rows - 1);
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    float u, v;
    map_xy(m, (float)((k & 1) ? cx1 : cx0), (float)((k & 2) ? cy1 : cy0), u, v);
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  // safe interior: every pixel of the tile (and of its 2-px erosion halo) is valid and every tap is in bounds
  if (!(umin >= 3.f && vmin >= 3.f && umax <= (float)(a.src_cols - 4) && vmax <= (float)(a.src_rows - 4))) return p;
  // footprint of the tile proper (the halo only served the safety test)
  umin = vmin = 3.4e38f; umax = vmax = -3.4e38f;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    float u, v;
    map_xy(m, (float)((k & 1) ? bx0 + TW - 1 : bx0), (float)((k & 2) ? by0 + TH - 1 : by0), u, v);
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  const int x_lo = (int)floorf(umin) - 1, x_hi = (int)floorf(umax) + 3;
  const int y_lo = (int)floorf(vmin) - 1, y_hi = (int)floorf(vmax) + 3;
  p.sx0 = x_lo & ~(align - 1);
  p.sy0 = y_lo;
  p.sxw = x_lo & ~3;
  p.staged = (x_hi - p.sx0 < wd) && (y_hi - p.sy0 < GSH) && (x_hi - p.sxw < WWD);
  return p;
}

template <int DEPTH>
__device__ __noinline__ void issue_stage(const FrameJob &job, const StagePlan &p, const WarpAccArgs &a, bool weighted,
                                         unsigned char *s_f, float *s_g) {
  typedef StageGeom<DEPTH> G;
  constexpr int CPR = G::WD / G::ALIGN;            // 16-byte chunks per staged frame row
  // one 64-bit base per tile; chunk offsets stay in 32 bits (a staged window spans GSH rows)
  const char *fbase = static_cast<const char *>(job.frame) + (int64_t)p.sy0 * a.src_step + (int64_t)p.sx0 * G::ES;
  const unsigned sf = (unsigned)__cvta_generic_to_shared(s_f);
  const int rows_left = a.src_rows - p.sy0, chunks_left = (a.src_cols - p.sx0) / G::ALIGN;   // in-bounds rows / whole chunks
  const int step = (int)a.src_step;
#pragma unroll
  for (int i = 0; i < (GSH * CPR + TW * SWARPS - 1) / (TW * SWARPS); ++i) {
    const int k = threadIdx.x + i * TW * SWARPS;
    const int r = k / CPR, q = k - r * CPR;
    if (k < GSH * CPR && r < rows_left && q < chunks_left)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sf + r * G::ROWB + q * 16), "l"(fbase + (r * step + q * 16)));
  }
  if (weighted) {
    const char *wbase = reinterpret_cast<const char *>(job.weights) + (int64_t)p.sy0 * a.w_step + (int64_t)p.sxw * 4;
    const unsigned sg = (unsigned)__cvta_generic_to_shared(s_g);
    const int wchunks_left = (a.src_cols - p.sxw) / 4;
    const int wstep = (int)a.w_step;
#pragma unroll
    for (int i = 0; i < (GSH * (WWD / 4) + TW * SWARPS - 1) / (TW * SWARPS); ++i) {
      const int k = threadIdx.x + i * TW * SWARPS;
      const int r = k / (WWD / 4), q = k - r * (WWD / 4);
      if (k < GSH * (WWD / 4) && r < rows_left && q < wchunks_left)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sg + (r * WWD + q * 4) * 4), "l"(wbase + (r * wstep + q * 16)));
    }
  }
}

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

// Staging of a border-ring tile: the (unclipped) footprint is materialised in shared memory as a virtually padded
// image - in-range 16-byte chunks by cp.async like the interior tiles, the overhang element-wise through
// cv::borderInterpolate (frame) or as zeros (weights: remapped with BORDER_CONSTANT 0).  The interpolation code is
// then the interior one.
template <int DEPTH>
__device__ __noinline__ void issue_stage_ring(const FrameJob &job, const StagePlan &p, const WarpAccArgs &a, bool weighted,
                                              unsigned char *s_f, float *s_g) {
  typedef StageGeom<DEPTH> G;
  typedef typename PixT<DEPTH>::type T;
  constexpr int CPR = G::WD / G::ALIGN;
#pragma unroll 1
  for (int k = threadIdx.x; k < GSH * CPR; k += blockDim.x) {
    const int r = k / CPR, q = k - r * CPR;
    const int gy = p.sy0 + r, gx = p.sx0 + q * G::ALIGN;
    unsigned char *d = s_f + r * G::ROWB + q * 16;
    if ((unsigned)gy < (unsigned)a.src_rows && gx >= 0 && gx + G::ALIGN <= a.src_cols) {
      cp_async16(d, static_cast<const char *>(job.frame) + (int64_t)gy * a.src_step + (int64_t)gx * G::ES);
    } else {
      // overhang: each element comes from its cv::borderInterpolate position (asynchronously, like the in-range
      // chunks) or is the constant border value
      const int my = bmap(gy, a.src_rows, a.border);
      const char *row = static_cast<const char *>(job.frame) + (int64_t)max(my, 0) * a.src_step;
#pragma unroll
      for (int e = 0; e < G::ALIGN; ++e) {
        const int mx = bmap(gx + e, a.src_cols, a.border);
        if (mx >= 0 && my >= 0) {
          if (G::ES == 4) cp_async4(d + e * 4, row + (int64_t)mx * 4);
          else reinterpret_cast<T *>(d)[e] = reinterpret_cast<const T *>(row)[mx];
        } else {
          reinterpret_cast<T *>(d)[e] = DEPTH == SSK_32F ? (T)a.bval[0] : (T)0;   // integer frames: staged only for value 0
        }
      }
    }
  }
  if (weighted) {
#pragma unroll 1
    for (int k = threadIdx.x; k < GSH * (WWD / 4); k += blockDim.x) {
      const int r = k / (WWD / 4), q = k - r * (WWD / 4);
      const int gy = p.sy0 + r, gx = p.sxw + q * 4;
      float *d = s_g + r * WWD + q * 4;
      if ((unsigned)gy < (unsigned)a.src_rows && gx >= 0 && gx + 4 <= a.src_cols) {
        cp_async16(d, reinterpret_cast<const char *>(job.weights) + (int64_t)gy * a.w_step + (int64_t)gx * 4);
      } else {
        const bool yin = (unsigned)gy < (unsigned)a.src_rows;
        const char *row = reinterpret_cast<const char *>(job.weights) + (int64_t)(yin ? gy : 0) * a.w_step;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (yin && (unsigned)(gx + e) < (unsigned)a.src_cols) cp_async4(d + e, row + (int64_t)(gx + e) * 4);
          else d[e] = 0.f;
        }
      }
    }
  }
}

// Staging plan of a border-ring tile: like plan_stage but without the in-bounds requirement.  Not staged (generic
// path) for BORDER_WRAP, projective maps, oversize footprints, overhang beyond one reflection, and a non-zero
// constant border on integer frames.
__device__ __noinline__ StagePlan plan_stage_ring(const MapCoef &m, int bx0, int by0, const WarpAccArgs &a, int align, int wd) {
  StagePlan p; p.staged = 0; p.sx0 = p.sy0 = p.sxw = 0;
  if (!a.stage_aligned || !is_affine_like(m.type) || a.border == SSK_BORDER_WRAP) return p;
  const bool constant = a.border != SSK_BORDER_REPLICATE && a.border != SSK_BORDER_REFLECT && a.border != SSK_BORDER_REFLECT101;
  if (constant && a.depth != SSK_32F && a.bval[0] != 0.f) return p;
  const int cx1 = min(bx0 + TW - 1, a.cols - 1), cy1 = min(by0 + TH - 1, a.rows - 1);
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    float u, v;
    map_xy(m, (float)((k & 1) ? cx1 : bx0), (float)((k & 2) ? cy1 : by0), u, v);
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  if (!(umax - umin < 64.f && vmax - vmin < 64.f)) return p;
  if (!(umin > -1.0e6f && vmin > -1.0e6f && umax < 1.0e6f && vmax < 1.0e6f)) return p;
  const int x_lo = (int)floorf(umin) - 1, x_hi = (int)floorf(umax) + 3;
  const int y_lo = (int)floorf(vmin) - 1, y_hi = (int)floorf(vmax) + 3;
  p.sx0 = x_lo & ~(align - 1);
  p.sy0 = y_lo;
  p.sxw = x_lo & ~3;
  // a single reflection must land inside the frame for every staged position
  const int lo_x = min(p.sx0, p.sxw), hi_x = max(p.sx0 + wd, p.sxw + WWD), hi_y = p.sy0 + GSH;
  if (-lo_x >= a.src_cols || hi_x - a.src_cols >= a.src_cols || -p.sy0 >= a.src_rows || hi_y - a.src_rows >= a.src_rows) return p;
  p.staged = (x_hi - p.sx0 < wd) && (y_hi - p.sy0 < GSH) && (x_hi - p.sxw < WWD);
  if (p.staged) {
    // staged = 2: every pixel of the tile and of its 2-px erosion halo (clipped to the image) maps into the tap-safe
    // interior of the frame, so the eroded validity mask is all ones for this frame and the flag pass is skipped
    const int hx0 = max(bx0 - 2, 0), hy0 = max(by0 - 2, 0), hx1 = min(bx0 + TW + 1, a.cols - 1), hy1 = min(by0 + TH + 1, a.rows - 1);
    float hu0 = 3.4e38f, hu1 = -3.4e38f, hv0 = 3.4e38f, hv1 = -3.4e38f;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      float u, v;
      map_xy(m, (float)((k & 1) ? hx1 : hx0), (float)((k & 2) ? hy1 : hy0), u, v);
      hu0 = fminf(hu0, u); hu1 = fmaxf(hu1, u); hv0 = fminf(hv0, v); hv1 = fmaxf(hv1, v);
    }
    if (hu0 >= 3.f && hv0 >= 3.f && hu1 <= (float)(a.src_cols - 4) && hv1 <= (float)(a.src_rows - 4)) p.staged = 2;
  }
  return p;
}

template <int DEPTH, int INTERP, bool WEIGHTS, int MT, bool RING>
__global__ void __launch_bounds__(TW * (RING ? RING_WARPS : SWARPS), RING ? SSK_RING_MINB : 6) k_fused_staged(const __grid_constant__ WarpAccArgs a, const __grid_constant__ Tables tab,
                                                               const TileList tl) {
  typedef StageGeom<DEPTH> G;
  constexpr int N = Taps<INTERP>::N;
  constexpr int NWARP = RING ? RING_WARPS : SWARPS;   // warps per CTA
  constexpr int GR = TH / NWARP;                      // rows per warp strip
  __shared__ float s_acc[TH][TW];                  // running mean of the tile (on chip for the whole batch)
  __shared__ float s_w[TH][TW];                    // running weight sum of the tile
  __shared__ float4 s_cubic[kInterTab];
  __shared__ __align__(128) unsigned char s_f[2][GSH * G::ROWB];   // 128-byte alignment: TMA destination
  __shared__ __align__(128) float s_g[2][GSH * WWD];
  __shared__ __align__(8) unsigned long long s_mbar[2];           // TMA path: one transaction barrier per buffer
  // ring tiles: per row of the tile + 2-px halo, the horizontally eroded validity of the tile columns (bit x = AND of
  // the pre-erosion flags of columns x-2 .. x+2)
  __shared__ unsigned long long s_hmask[RING ? TH + 4 : 1];
  __shared__ PackedPlan s_plan[KPLAN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tx, ty;
  if (RING) tile_of_block(tl, blockIdx.x, tx, ty);
  else { tx = blockIdx.x + 1; ty = blockIdx.y + 1; }
  const int bx0 = tx * TW, by0 = ty * TH;
  const int x = bx0 + lane, y0 = by0 + warp * GR;
  const int tw = min(TW, a.cols - bx0), th = min(TH, a.rows - by0);
  const int nrow = lane < tw ? max(0, min(GR, th - warp * GR)) : 0;
  if (INTERP == SSK_INTER_CUBIC && threadIdx.x < kInterTab) s_cubic[threadIdx.x] = tab.cubic[threadIdx.x];

  // accumulator tile -> shared memory; 16-byte accesses when the tile is complete and the pitch allows it
  const bool vec = (a.cols & 3) == 0 && tw == TW;
  if (vec) {
    for (int k = threadIdx.x; k < th * (TW / 4); k += blockDim.x) {
      const int r = k / (TW / 4), q = k - r * (TW / 4);
      reinterpret_cast<float4 *>(s_acc[r])[q] = *reinterpret_cast<const float4 *>(a.acc + (int64_t)(by0 + r) * a.cols + bx0 + 4 * q);
      reinterpret_cast<float4 *>(s_w[r])[q] = *reinterpret_cast<const float4 *>(a.wacc + (int64_t)(by0 + r) * a.cols + bx0 + 4 * q);
    }
  } else {
    for (int k = threadIdx.x; k < th * tw; k += blockDim.x) {
      const int r = k / tw, q = k - r * tw;
      s_acc[r][q] = a.acc[(int64_t)(by0 + r) * a.cols + bx0 + q];
      s_w[r][q] = a.wacc[(int64_t)(by0 + r) * a.cols + bx0 + q];
    }
  }

  // staging plans of every frame of the launch, one frame per thread
  for (int jj = threadIdx.x; jj < a.njobs; jj += blockDim.x) {
    PackedPlan pp = {0, 0, 0, -1};
    if (a.jobs[jj].ok) {
      StagePlan p = RING ? plan_stage_ring(a.jobs[jj].map, bx0, by0, a, G::ALIGN, G::WD) : plan_stage(a.jobs[jj].map, bx0, by0, a, G::ALIGN, G::WD);
      if (WEIGHTS && !a.jobs[jj].weights) p.staged = 0;    // flat frame (no weight map): generic path
      pp.sx0 = (short)p.sx0; pp.sy0 = (short)p.sy0; pp.sxw = (short)p.sxw; pp.staged = (short)p.staged;
    }
    s_plan[jj] = pp;
  }
  // TMA staging: interior tiles of 32F frames with weight maps when the host supplied tensor maps (ssk_stack.cu)
  const bool use_tma = !RING && DEPTH == SSK_32F && WEIGHTS && a.tmap_frames != nullptr && a.tmap_weights != nullptr;
  const unsigned mbar0 = (unsigned)__cvta_generic_to_shared(&s_mbar[0]);
  if (use_tma && threadIdx.x == 0) {
    mbar_init(mbar0, 1);
    mbar_init(mbar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned phase = 0;   // bit b: parity the next wait on buffer b expects
  auto tma_stage = [&](int jj, const StagePlan &p, int b) {
    if (threadIdx.x == 0) {
      const char *tf = static_cast<const char *>(a.tmap_frames) + (size_t)jj * 128, *tw = static_cast<const char *>(a.tmap_weights) + (size_t)jj * 128;
      tmap_acquire(tf);
      tmap_acquire(tw);
      const unsigned mb = mbar0 + 8 * b;
      mbar_expect_tx(mb, (unsigned)(GSH * G::ROWB + GSH * WWD * 4));
      tma_load_2d((unsigned)__cvta_generic_to_shared(s_f[b]), tf, p.sx0, p.sy0, mb);
      tma_load_2d((unsigned)__cvta_generic_to_shared(s_g[b]), tw, p.sxw, p.sy0, mb);
    }
  };
  __syncthreads();
  auto plan_of = [&](int jj) { const PackedPlan q = s_plan[jj]; StagePlan p; p.staged = q.staged; p.sx0 = q.sx0; p.sy0 = q.sy0; p.sxw = q.sxw; return p; };

  // software pipeline over the frames of the batch: while frame j is interpolated, frame j+1's footprint lands
  int j = 0;
  while (j < a.njobs && s_plan[j].staged < 0) ++j;
  int buf = 0;
  StagePlan plan = {0, 0, 0, 0};
  if (j < a.njobs) {
    plan = plan_of(j);
    if (plan.staged) {
      if (RING) issue_stage_ring<DEPTH>(a.jobs[j], plan, a, WEIGHTS, s_f[0], s_g[0]);
      else if (use_tma) tma_stage(j, plan, 0);
      else issue_stage<DEPTH>(a.jobs[j], plan, a, WEIGHTS, s_f[0], s_g[0]);
    }
  }
  cp_async_commit();
  __syncthreads();

#pragma unroll 1
  while (j < a.njobs) {
    int jn = j + 1;
    while (jn < a.njobs && s_plan[jn].staged < 0) ++jn;
    StagePlan plan_n = {0, 0, 0, 0};
    if (jn < a.njobs) {
      plan_n = plan_of(jn);
      if (plan_n.staged) {
        if (RING) issue_stage_ring<DEPTH>(a.jobs[jn], plan_n, a, WEIGHTS, s_f[buf ^ 1], s_g[buf ^ 1]);
        else if (use_tma) tma_stage(jn, plan_n, buf ^ 1);
        else issue_stage<DEPTH>(a.jobs[jn], plan_n, a, WEIGHTS, s_f[buf ^ 1], s_g[buf ^ 1]);
      }
    }
    cp_async_commit();
    if (RING && plan.staged == 1) {
      // pre-erosion validity of the tile and its 2-px halo (outside the image: erode border value 255), one warp
      // per row, packed by ballots
      // columns -2 .. 29 of a row go through one ballot; columns 30 .. 33 of all rows of this warp share one more
      const MapCoef m = a.jobs[j].map;
      constexpr int NR = (TH + 4 + NWARP - 1) / NWARP;     // rows of the flag field per warp
      static_assert(!RING || NR <= 8, "ring flags: one nibble per row in the second ballot");
      unsigned lo[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int gy = by0 - 2 + warp + i * NWARP, gx = bx0 - 2 + lane;
        bool okf = true;
        if (warp + i * NWARP < TH + 4 && gx >= 0 && gy >= 0 && gx < a.cols && gy < a.rows) {
          float u, v;
          const ColMap<MT> cmf(m, (float)gx);
          cmf((float)gy, u, v);
          okf = valid255(INTERP, u, v, a.src_cols, a.src_rows, tab.cubic_itab);
        }
        lo[i] = __ballot_sync(0xffffffffu, okf);
      }
      unsigned hi;
      {
        const int i = lane >> 2, fy = warp + i * NWARP;
        const int gy = by0 - 2 + fy, gx = bx0 + 30 + (lane & 3);
        bool okf = true;
        if (i < NR && fy < TH + 4 && gy >= 0 && gx < a.cols && gy < a.rows) {
          float u, v;
          const ColMap<MT> cmf(m, (float)gx);
          cmf((float)gy, u, v);
          okf = valid255(INTERP, u, v, a.src_cols, a.src_rows, tab.cubic_itab);
        }
        hi = __ballot_sync(0xffffffffu, okf);
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int fy = warp + i * NWARP;
        if (lane == 0 && fy < TH + 4) {
          const unsigned long long bits = (unsigned long long)lo[i] | ((unsigned long long)((hi >> (4 * i)) & 0xFu) << 32);
          s_hmask[fy] = bits & (bits >> 1) & (bits >> 2) & (bits >> 3) & (bits >> 4);
        }
      }
    }
    if (use_tma) {
      // the copy engine signals the buffer's barrier when both windows of frame j have landed: no CTA barrier needed
      if (plan.staged) { mbar_wait(mbar0 + 8 * buf, (phase >> buf) & 1u); phase ^= 1u << buf; }
    } else {
      cp_async_wait<1>();          // frame j's group has landed (frame j+1's may still be in flight)
      __syncthreads();
    }

    float *s_acc0 = &s_acc[warp * GR][lane], *s_w0 = &s_w[warp * GR][lane];
    if (plan.staged) {
      const MapCoef m = a.jobs[j].map;
      const ColMap<MT> cm(m, (float)x);
      const unsigned char *sf = s_f[buf];
      const float *sg = s_g[buf];
      constexpr bool C2 = INTERP == SSK_INTER_CUBIC && WEIGHTS;   // packed (frame, weight) bicubic
      const unsigned acc_a = (unsigned)__cvta_generic_to_shared(s_acc0), w_a = (unsigned)__cvta_generic_to_shared(s_w0);
      const unsigned cub_a = (unsigned)__cvta_generic_to_shared(s_cubic);
      const float y0f = (float)y0;
      if (RING) {
        RollS<INTERP> R;
        RollC2 R2;
        R.ix = INT_MIN; R.iy = INT_MIN; R.pf = sf; R.pw = sg;
        R2.ix = INT_MIN; R2.iy = INT_MIN; R2.pf = sf; R2.pw = sg;
        const unsigned long long *hm = &s_hmask[warp * GR];
        unsigned long long m0 = hm[0], m1 = hm[1], m2 = hm[2], m3 = hm[3];
#pragma unroll 1
        for (int k = 0; k < GR; k += N) {
#pragma unroll
          for (int jj = 0; jj < N; ++jj) {
            const unsigned long long m4 = k + jj < GR ? hm[k + jj + 4] : 0ull;
            const bool ok = plan.staged == 2 || (((m0 & m1 & m2 & m3 & m4) >> lane) & 1ull);
            m0 = m1; m1 = m2; m2 = m3; m3 = m4;
            if (k + jj < nrow) {
              if (C2) {
                const unsigned o = (unsigned)((k + jj) * TW * 4);
                if (jj == 0) roll_pixel_c2<DEPTH, MT, 0>(R2, cm, y0f + (float)k, sf, sg, a.scale, plan, cub_a, acc_a + o, w_a + o, ok);
                if (jj == 1) roll_pixel_c2<DEPTH, MT, 1>(R2, cm, y0f + (float)(k + 1), sf, sg, a.scale, plan, cub_a, acc_a + o, w_a + o, ok);
                if (jj == 2) roll_pixel_c2<DEPTH, MT, 2>(R2, cm, y0f + (float)(k + 2), sf, sg, a.scale, plan, cub_a, acc_a + o, w_a + o, ok);
                if (jj == 3) roll_pixel_c2<DEPTH, MT, 3>(R2, cm, y0f + (float)(k + 3), sf, sg, a.scale, plan, cub_a, acc_a + o, w_a + o, ok);
              } else {
                if (jj == 0) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 0>(R, cm, (float)(y0 + k), sf, sg, a.scale, plan, s_cubic, s_acc0 + k * TW, s_w0 + k * TW, ok);
                if (jj == 1) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 1 % N>(R, cm, (float)(y0 + k + 1), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 1) * TW, s_w0 + (k + 1) * TW, ok);
                if (jj == 2) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 2 % N>(R, cm, (float)(y0 + k + 2), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 2) * TW, s_w0 + (k + 2) * TW, ok);
                if (jj == 3) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 3 % N>(R, cm, (float)(y0 + k + 3), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 3) * TW, s_w0 + (k + 3) * TW, ok);
              }
            }
          }
        }
      } else if (C2) {
        RollC2 R2;
        R2.ix = INT_MIN; R2.iy = INT_MIN; R2.pf = sf; R2.pw = sg;
#pragma unroll
        for (int k = 0; k < GR; k += 4) {
          roll_pixel_c2<DEPTH, MT, 0>(R2, cm, y0f + (float)k, sf, sg, a.scale, plan, cub_a, acc_a + k * TW * 4, w_a + k * TW * 4);
          roll_pixel_c2<DEPTH, MT, 1>(R2, cm, y0f + (float)(k + 1), sf, sg, a.scale, plan, cub_a, acc_a + (k + 1) * TW * 4, w_a + (k + 1) * TW * 4);
          roll_pixel_c2<DEPTH, MT, 2>(R2, cm, y0f + (float)(k + 2), sf, sg, a.scale, plan, cub_a, acc_a + (k + 2) * TW * 4, w_a + (k + 2) * TW * 4);
          roll_pixel_c2<DEPTH, MT, 3>(R2, cm, y0f + (float)(k + 3), sf, sg, a.scale, plan, cub_a, acc_a + (k + 3) * TW * 4, w_a + (k + 3) * TW * 4);
        }
      } else {
        RollS<INTERP> R;
        R.ix = INT_MIN; R.iy = INT_MIN; R.pf = sf; R.pw = sg;
#pragma unroll 1
        for (int k = 0; k < GR; k += N) {
          roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 0>(R, cm, (float)(y0 + k), sf, sg, a.scale, plan, s_cubic, s_acc0 + k * TW, s_w0 + k * TW);
          if (N > 1) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 1 % N>(R, cm, (float)(y0 + k + 1), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 1) * TW, s_w0 + (k + 1) * TW);
          if (N > 2) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 2 % N>(R, cm, (float)(y0 + k + 2), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 2) * TW, s_w0 + (k + 2) * TW);
          if (N > 3) roll_pixel_s<DEPTH, INTERP, WEIGHTS, MT, 3 % N>(R, cm, (float)(y0 + k + 3), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 3) * TW, s_w0 + (k + 3) * TW);
        }
      }
    } else {
      // this (tile, frame) pair cannot be staged (large displacement, flat frame, BORDER_WRAP): generic per-pixel path
#pragma unroll 1
      for (int k = 0; k < nrow; ++k) generic_pixel(a, tab, a.jobs[j], x, y0 + k, s_acc0 + k * TW, s_w0 + k * TW);
    }
    __syncthreads();               // everyone is done with buffer `buf` (and the flags) before they are refilled
    j = jn; plan = plan_n; buf ^= 1;
  }
  cp_async_wait<0>();
  __syncthreads();

  if (vec) {
    for (int k = threadIdx.x; k < th * (TW / 4); k += blockDim.x) {
      const int r = k / (TW / 4), q = k - r * (TW / 4);
      *reinterpret_cast<float4 *>(a.acc + (int64_t)(by0 + r) * a.cols + bx0 + 4 * q) = reinterpret_cast<const float4 *>(s_acc[r])[q];
      *reinterpret_cast<float4 *>(a.wacc + (int64_t)(by0 + r) * a.cols + bx0 + 4 * q) = reinterpret_cast<const float4 *>(s_w[r])[q];
    }
  } else {
    for (int k = threadIdx.x; k < th * tw; k += blockDim.x) {
      const int r = k / tw, q = k - r * tw;
      a.acc[(int64_t)(by0 + r) * a.cols + bx0 + q] = s_acc[r][q];
      a.wacc[(int64_t)(by0 + r) * a.cols + bx0 + q] = s_w[r][q];
    }
  }
}

template <int DEPTH, int INTERP, bool WEIGHTS, int MT>
void launch_staged_ring(const WarpAccArgs &a, const Tables &tab, const TileList &tl, int nring, cudaStream_t s) {
  const dim3 block(TW * SWARPS), ring_block(TW * RING_WARPS);
  // The border ring (few, heavier CTAs) runs on a side stream so that it overlaps the interior tiles; the two
  // kernels write disjoint accumulator tiles.
  cudaStream_t ring_stream = s;
  if (a.side_stream) {
    ring_stream = static_cast<cudaStream_t>(a.side_stream);
    cudaEventRecord(static_cast<cudaEvent_t>(a.ev_fork), s);
    cudaStreamWaitEvent(ring_stream, static_cast<cudaEvent_t>(a.ev_fork), 0);
  }
  k_fused_staged<DEPTH, INTERP, WEIGHTS, MT, true><<<nring, ring_block, 0, ring_stream>>>(a, tab, tl);
  count_launch();
  if (a.side_stream) cudaEventRecord(static_cast<cudaEvent_t>(a.ev_join), ring_stream);
  k_fused_staged<DEPTH, INTERP, WEIGHTS, MT, false><<<dim3(tl.ntx - 2, tl.nty - 2), block, 0, s>>>(a, tab, tl);
  if (a.side_stream && !a.defer_join) cudaStreamWaitEvent(s, static_cast<cudaEvent_t>(a.ev_join), 0);
}

template <int DEPTH, int INTERP, bool WEIGHTS>
void launch_staged_mt(const WarpAccArgs &a, const Tables &tab, const TileList &tl, int nring, cudaStream_t s) {
  if (a.map_type == MAP_AFFINE) launch_staged_ring<DEPTH, INTERP, WEIGHTS, MAP_AFFINE>(a, tab, tl, nring, s);
  else if (a.map_type == MAP_TRANSLATION) launch_staged_ring<DEPTH, INTERP, WEIGHTS, MAP_TRANSLATION>(a, tab, tl, nring, s);
  else launch_staged_ring<DEPTH, INTERP, WEIGHTS, MAP_EUCLIDEAN>(a, tab, tl, nring, s);
}

template <int DEPTH>
void launch_staged(const WarpAccArgs &a, const Tables &tab, const TileList &tl, int nring, cudaStream_t s) {
  if (a.use_weights) {
    if (a.interp == SSK_INTER_CUBIC) launch_staged_mt<DEPTH, SSK_INTER_CUBIC, true>(a, tab, tl, nring, s);
    else if (a.interp == SSK_INTER_NEAREST) launch_staged_mt<DEPTH, SSK_INTER_NEAREST, true>(a, tab, tl, nring, s);
    else launch_staged_mt<DEPTH, SSK_INTER_LINEAR, true>(a, tab, tl, nring, s);
  } else {
    if (a.interp == SSK_INTER_CUBIC) launch_staged_mt<DEPTH, SSK_INTER_CUBIC, false>(a, tab, tl, nring, s);
    else if (a.interp == SSK_INTER_NEAREST) launch_staged_mt<DEPTH, SSK_INTER_NEAREST, false>(a, tab, tl, nring, s);
    else launch_staged_mt<DEPTH, SSK_INTER_LINEAR, false>(a, tab, tl, nring, s);
  }
}

}  // namespace

int staged_box_w() { return StageGeom<SSK_32F>::WD; }
int staged_box_h() { return GSH; }

int launch_warp_accumulate(const WarpAccArgs &a_in, const Tables &tab, cudaStream_t s) {
  WarpAccArgs a = a_in;
  SSK_REQUIRE(a.interp == SSK_INTER_NEAREST || a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC,
              "warp_accumulate: interpolation must be NEAREST, LINEAR or CUBIC");
  SSK_REQUIRE(a.border != SSK_BORDER_TRANSPARENT, "warp_accumulate: BORDER_TRANSPARENT is not meaningful here");
  SSK_REQUIRE(a.cn >= 1 && a.cn <= 4, "warp_accumulate: 1..4 channels");
  SSK_REQUIRE(a.depth == SSK_32F || a.depth == SSK_16U || a.depth == SSK_8U, "warp_accumulate: unsupported frame depth");
  a.stage_aligned = a.stage_aligned && (a.src_step % 16 == 0) && (a.w_step % 16 == 0);
  const int ntx = div_up(a.cols, TW), nty = div_up(a.rows, TH);
  const bool staged = a.cn == 1 && ntx >= 3 && nty >= 3 && a.src_cols < 32000 && a.src_rows < 32000 &&
                      (a.map_type == MAP_AFFINE || a.map_type == MAP_TRANSLATION || a.map_type == MAP_EUCLIDEAN);
  TileList tl;
  tl.ntx = ntx; tl.nty = nty; tl.ring = staged ? 1 : 0;
  if (staged) {
    const int nring = 2 * ntx + 2 * (nty - 2);
    const FrameJob *jobs = a.jobs;
    const int njobs = a.njobs;
    for (int j0 = 0; j0 < njobs; j0 += KPLAN) {     // the per-CTA plan table holds KPLAN frames
      a.jobs = jobs + j0; a.njobs = std::min(KPLAN, njobs - j0);
      if (a_in.tmap_frames && a_in.tmap_weights) {
        a.tmap_frames = static_cast<const char *>(a_in.tmap_frames) + (size_t)j0 * 128;
        a.tmap_weights = static_cast<const char *>(a_in.tmap_weights) + (size_t)j0 * 128;
      } else {
        a.tmap_frames = a.tmap_weights = nullptr;
      }
      if (a.depth == SSK_32F) launch_staged<SSK_32F>(a, tab, tl, nring, s);
      else if (a.depth == SSK_16U) launch_staged<SSK_16U>(a, tab, tl, nring, s);
      else launch_staged<SSK_8U>(a, tab, tl, nring, s);
      SSK_LAUNCH_CHECK();
    }
  } else {
    // multi-channel frames, projective maps, tiny images: one thread per pixel
    k_fused_generic<<<ntx * nty * 4, 256, 0, s>>>(a, tab, tl);
    SSK_LAUNCH_CHECK();
  }
  return SSK_OK;
}

}  // namespace ssk
