// c_eccflow on the device: dense smooth optical flow, batched over the frames of a stacking batch.
//
// Reference semantics (core/proc/image_registration/):
//   c_eccflow::set_reference_image   ecc2.cc:2494-2672   pyramid (INTER_AREA recursion / full resize / pyrDown), Ix, Iy,
//                                                        D = (<Ix Ix>, <Ix Iy>, <Iy Iy>, update_multiplier / (|det| + reg))
//   c_eccflow::setup_input_image     ecc2.cc:2674-2768   current pyramid with the same sizes
//   c_eccflow::compute_uv (level)    ecc2.cc:2237-2398   W = remap(current, uv + grid, LINEAR, REPLICATE); It = ref - W;
//                                                        (It Ix, It Iy) -> avgdown -> 2x2 solve with D -> resize(CUBIC)
//   c_eccflow::avgdown               ecc2.cc:2400-2418   resize(INTER_AREA) to the support_scale-times halved size, then
//                                                        sepFilter2D with getGaussianKernel(3, 0) = (1/4, 1/2, 1/4), REPLICATE
//   c_eccflow::compute_uv / compute  ecc2.cc:2773-2865   initial flow from the map, coarse-to-fine, max_iterations per level
//
// Device structure.  The reference forms five full-size intermediates per iteration (W, M, It, Itxy, the resized update);
// here an iteration of a level is three launches for ALL frames of the batch:
//   k_flow_reduce   one CTA per (coarse row, frame): walks the source rows of that coarse row, samples the current image
//                   through the flow (cv::remap's 1/32-px bilinear, bit-exact), forms It Ix / It Iy in registers and
//                   reduces them over the INTER_AREA cells of the row (column sums in shared memory): the only full-size
//                   traffic is one read of uv, ref, Ix, Iy and the bilinear gather of the current image
//   k_flow_solve    3 x 3 Gaussian of the coarse sums + the 2 x 2 solve (coarse grid: 1/256 of the pixels)
//   k_flow_update   uv += resize(update, INTER_CUBIC) evaluated per pixel from the coarse grid (L1/L2-resident)
// Level changes (uv * size ratio -> resize CUBIC) and the initial flow (map - grid -> resize CUBIC -> * ratio) are one
// launch each; an analytic registration map (the ECC result) is evaluated on the fly, it is never materialised.
// Sums are formed in a different order than OpenCV forms them (column-first instead of row-first), so the flow agrees with
// the reference to float rounding, not bit for bit; tests/test_gpu_eccflow.py states the tolerance.
#include <cmath>
#include <cstring>
#include <algorithm>
#include "ssk_eccflow.cuh"

namespace ssk {

namespace {

constexpr float kD5[5] = {1.0f / 12.0f, -2.0f / 3.0f, 0.0f, 2.0f / 3.0f, -1.0f / 12.0f};
constexpr float kS3[3] = {0.25f, 0.5f, 0.25f};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// cv::resize(INTER_CUBIC) of a 2-channel float image at one destination pixel: horizontal pass of the four source rows, then
// the vertical combination (HResizeCubic / VResizeCubic), taps clamped into the image
template <class LD>
__device__ __forceinline__ float2 cubic_at(LD ld, int sw, int sh, int sx, const float4 cx, int sy, const float4 cy) {
  const int x0 = clampi(sx - 1, 0, sw - 1), x1 = clampi(sx, 0, sw - 1), x2 = clampi(sx + 1, 0, sw - 1), x3 = clampi(sx + 2, 0, sw - 1);
  float2 r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = clampi(sy - 1 + k, 0, sh - 1);
    const float2 a0 = ld(x0, yy), a1 = ld(x1, yy), a2 = ld(x2, yy), a3 = ld(x3, yy);
    r[k].x = a0.x * cx.x + a1.x * cx.y + a2.x * cx.z + a3.x * cx.w;
    r[k].y = a0.y * cx.x + a1.y * cx.y + a2.y * cx.z + a3.y * cx.w;
  }
  float2 o;
  o.x = r[0].x * cy.x + r[1].x * cy.y + r[2].x * cy.z + r[3].x * cy.w;
  o.y = r[0].y * cy.x + r[1].y * cy.y + r[2].y * cy.z + r[3].y * cy.w;
  return o;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_flow_reduce
// ---------------------------------------------------------------------------------------------------------------------
// cv::remap's bilinear sample (sample_linear of ssk_common.cuh with BORDER_REPLICATE) split in two steps, so that a thread can
// have the taps of several pixels in flight: offsets + weights first, the arithmetic after the loads
struct LinTaps { int o00, o01, o10, o11; float w00, w01, w10, w11; };
// read-only loads the compiler keeps in program order relative to each other (asm volatile): the batched loop below wants all
// loads of a step issued before the first use, ptxas otherwise re-serialises them to save registers
__device__ __forceinline__ float ldg_ord(const float *p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_ord2(const float2 *p) {
  float2 v;
  asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ LinTaps lin_taps(float u, float v, int w, int h) {
  int ix, fx, iy, fy;
  quant32(u, ix, fx);
  quant32(v, iy, fy);
  const float tx = (float)fx * 0.03125f, ty = (float)fy * 0.03125f;
  const float wx0 = 1.0f - tx, wy0 = 1.0f - ty;
  const int x0 = clampi(ix, 0, w - 1), x1 = clampi(ix + 1, 0, w - 1), y0 = clampi(iy, 0, h - 1), y1 = clampi(iy + 1, 0, h - 1);
  LinTaps t;
  t.o00 = y0 * w + x0; t.o01 = y0 * w + x1; t.o10 = y1 * w + x0; t.o11 = y1 * w + x1;
  t.w00 = __fmul_rn(wy0, wx0); t.w01 = __fmul_rn(wy0, tx); t.w10 = __fmul_rn(ty, wx0); t.w11 = __fmul_rn(ty, tx);
  return t;
}
__device__ __forceinline__ float lin_combine(const LinTaps &t, float s00, float s01, float s10, float s11) {
  float out = __fadd_rn(__fmul_rn(s00, t.w00), __fmul_rn(s01, t.w01));
  out = __fadd_rn(out, __fmul_rn(s10, t.w10));
  return __fadd_rn(out, __fmul_rn(s11, t.w11));
}

struct FlowReduceArgs {
  int w, h, cw, ch;
  const float *ref, *ix, *iy;            // level images (dense)
  const uint8_t *refmask;                // level mask or null
  const float *cur; int64_t cur_stride;  // frame b: cur + b * cur_stride (floats)
  const uint8_t *curmask;                // level mask of frame 0 (batch == 1) or null
  const float2 *uv; int64_t uv_stride;   // frame b: uv + b * uv_stride (float2)
  // the update of the previous iteration of this level applied on the way instead of by a k_flow_update pass of its own:
  // f = uv + resize(cuv_prev, INTER_CUBIC) is used and written to uv_out (another buffer: a source row shared by two coarse
  // rows is read by both CTAs and written by the first one only).  null: uv is used as it is
  const float2 *cuv_prev; int64_t cuv_stride;
  float2 *uv_out;
  FlowCubicAxis ux, uy;
  float *out; int64_t out_stride;        // frame b: raw coarse sums [ch][cw][NOUT]
  FlowAreaAxis ax, ay;
  const EccFrame *frames;                // ok flags of the batch or null
};

// MODE 0: (It Ix, It Iy) of the flow iteration; MODE 1: (Ix Ix, Ix Iy, Iy Iy) of the reference side (avgp)
// FUSE (MODE 0 only): apply the previous iteration's update on the way (a.cuv_prev, a.uv_out)
template <int MODE, int FUSE>
__global__ void __launch_bounds__(256) k_flow_reduce(const FlowReduceArgs a) {
  constexpr int NOUT = MODE == 0 ? 2 : 3;
  extern __shared__ float s_col[];       // [NOUT][w]
  const int cy = blockIdx.x, b = blockIdx.y;
  if (MODE == 0 && a.frames && !a.frames[b].ok) return;
  const int y0 = a.ay.start[cy], ny = a.ay.count[cy];
  const float *beta = a.ay.alpha + a.ay.off[cy];
  Img im;
  im.data = a.cur + (int64_t)b * a.cur_stride; im.step = (int64_t)a.w * 4; im.rows = a.h; im.cols = a.w;
  im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
  const float2 *__restrict__ uv = a.uv + (int64_t)b * a.uv_stride;
  const float2 *__restrict__ cprev = FUSE ? a.cuv_prev + (int64_t)b * a.cuv_stride : nullptr;
  float2 *__restrict__ uv_out = FUSE ? a.uv_out + (int64_t)b * a.uv_stride : nullptr;
  const int own_from = (FUSE && cy > 0) ? a.ay.start[cy - 1] + a.ay.count[cy - 1] : 0;   // rows below belong to the coarse row above
  for (int x = threadIdx.x; x < a.w; x += 256) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    int uxs = 0; float4 uxc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (FUSE) { uxs = __ldg(a.ux.s + x); uxc = __ldg(a.ux.c + x); }
    int kbeg = 0;
    if (MODE == 0 && !FUSE && !a.refmask && !a.curmask) {
      // four rows per step: all coalesced loads, then all 16 gathers, then the arithmetic (the loop below is latency-bound
      // with one pixel in flight per thread)
      const float *__restrict__ cur = static_cast<const float *>(im.data);
      for (; kbeg + 4 <= ny; kbeg += 4) {
        float2 f[4]; float gxs[4], gys[4], rf[4], bk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t p = (int64_t)(y0 + kbeg + j) * a.w + x;
          f[j] = ldg_ord2(uv + p);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t p = (int64_t)(y0 + kbeg + j) * a.w + x;
          gxs[j] = ldg_ord(a.ix + p); gys[j] = ldg_ord(a.iy + p); rf[j] = ldg_ord(a.ref + p); bk[j] = __ldg(beta + kbeg + j);
        }
        LinTaps t[4]; float s[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = lin_taps(__fadd_rn(f[j].x, (float)x), __fadd_rn(f[j].y, (float)(y0 + kbeg + j)), a.w, a.h);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[j][0] = ldg_ord(cur + t[j].o00); s[j][1] = ldg_ord(cur + t[j].o01); s[j][2] = ldg_ord(cur + t[j].o10); s[j][3] = ldg_ord(cur + t[j].o11);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float It = __fsub_rn(rf[j], lin_combine(t[j], s[j][0], s[j][1], s[j][2], s[j][3]));
          acc0 = fmaf(bk[j], __fmul_rn(It, gxs[j]), acc0);
          acc1 = fmaf(bk[j], __fmul_rn(It, gys[j]), acc1);
        }
      }
    }
#pragma unroll 1
    for (int k = kbeg; k < ny; ++k) {
      const int y = y0 + k;
      const int64_t p = (int64_t)y * a.w + x;
      const float gx = __ldg(a.ix + p), gy = __ldg(a.iy + p);
      const float bk = __ldg(beta + k);
      if (MODE == 0) {
        float2 f = __ldg(uv + p);
        if (FUSE) {
          const float2 d = cubic_at([&](int sx, int sy) { return __ldg(cprev + (int64_t)sy * a.cw + sx); }, a.cw, a.ch, uxs, uxc, __ldg(a.uy.s + y),
                                    __ldg(a.uy.c + y));
          f = make_float2(__fadd_rn(f.x, d.x), __fadd_rn(f.y, d.y));
          if (y >= own_from) uv_out[p] = f;
        }
        const float u = __fadd_rn(f.x, (float)x), v = __fadd_rn(f.y, (float)y);      // ecc_flow_to_remap
        bool ok = true;
        if (a.refmask) ok = __ldg(a.refmask + p) != 0;
        if (ok && a.curmask) {                                                        // remap(mask, INTER_NEAREST, CONSTANT 0)
          const int mx = __float2int_rn(u), my = __float2int_rn(v);
          ok = (unsigned)mx < (unsigned)a.w && (unsigned)my < (unsigned)a.h && __ldg(a.curmask + (int64_t)my * a.w + mx) != 0;
        }
        if (ok) {
          const float I1 = sample_linear<SSK_32F>(im, 0, u, v, SSK_BORDER_REPLICATE, 0.f);
          const float It = __fsub_rn(__ldg(a.ref + p), I1);
          acc0 = fmaf(bk, __fmul_rn(It, gx), acc0);
          acc1 = fmaf(bk, __fmul_rn(It, gy), acc1);
        }
      } else {
        acc0 = fmaf(bk, __fmul_rn(gx, gx), acc0);
        acc1 = fmaf(bk, __fmul_rn(gx, gy), acc1);
        acc2 = fmaf(bk, __fmul_rn(gy, gy), acc2);
      }
    }
    s_col[x] = acc0; s_col[a.w + x] = acc1;
    if (NOUT == 3) s_col[2 * a.w + x] = acc2;
  }
  __syncthreads();
  float *out = a.out + (int64_t)b * a.out_stride + (int64_t)cy * a.cw * NOUT;
  for (int cx = threadIdx.x; cx < a.cw; cx += 256) {
    const int x0 = a.ax.start[cx], nx = a.ax.count[cx];
    const float *alpha = a.ax.alpha + a.ax.off[cx];
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    for (int k = 0; k < nx; ++k) {
      const float al = __ldg(alpha + k);
      r0 = fmaf(al, s_col[x0 + k], r0);
      r1 = fmaf(al, s_col[a.w + x0 + k], r1);
      if (NOUT == 3) r2 = fmaf(al, s_col[2 * a.w + x0 + k], r2);
    }
    out[cx * NOUT] = r0; out[cx * NOUT + 1] = r1;
    if (NOUT == 3) out[cx * NOUT + 2] = r2;
  }
}

// sepFilter2D((1/4, 1/2, 1/4) x (1/4, 1/2, 1/4), BORDER_REPLICATE) of an interleaved NC-channel coarse image at (x, y):
// rows first, then columns (the products by 1/4 and 1/2 are exact, so every sum is rounded once whatever the fusing)
template <int NC>
__device__ __forceinline__ void gauss3(const float *src, int cw, int ch, int x, int y, float *out) {
  const int xm = max(x - 1, 0), xp = min(x + 1, cw - 1);
  float r[3][NC];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int yy = clampi(y - 1 + k, 0, ch - 1);
    const float *row = src + (int64_t)yy * cw * NC;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      r[k][c] = __fadd_rn(__fmul_rn(row[x * NC + c], 0.5f), __fmul_rn(__fadd_rn(row[xm * NC + c], row[xp * NC + c]), 0.25f));
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) out[c] = __fadd_rn(__fmul_rn(r[1][c], 0.5f), __fmul_rn(__fadd_rn(r[0][c], r[2][c]), 0.25f));
}

// reference side: D = (a00, a01, a11, update_multiplier / (|a00 a11 - a01 a01| + reg)) (ecc2.cc:2611-2650)
__global__ void __launch_bounds__(256) k_flow_D(const float *raw, int cw, int ch, float reg, float um, float4 *D) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= cw * ch) return;
  const int y = i / cw, x = i - y * cw;
  float a[3];
  gauss3<3>(raw, cw, ch, x, y, a);
  const float det = fabsf(__fsub_rn(__fmul_rn(a[0], a[2]), __fmul_rn(a[1], a[1])));
  D[i] = make_float4(a[0], a[1], a[2], __fdiv_rn(um, __fadd_rn(det, reg)));
}

// flow iteration: update on the coarse grid (ecc2.cc:2352-2388)
__global__ void __launch_bounds__(256) k_flow_solve(const float *raw, int64_t raw_stride, const float4 *D, int cw, int ch, float2 *cuv,
                                                    int64_t cuv_stride, const EccFrame *frames) {
  const int i = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (i >= cw * ch) return;
  if (frames && !frames[b].ok) return;
  const int y = i / cw, x = i - y * cw;
  float t[2];
  gauss3<2>(raw + (int64_t)b * raw_stride, cw, ch, x, y, t);
  const float4 d = __ldg(D + i);
  const float u = __fmul_rn(d.w, __fsub_rn(__fmul_rn(d.z, t[0]), __fmul_rn(d.y, t[1])));
  const float v = __fmul_rn(d.w, __fsub_rn(__fmul_rn(d.x, t[1]), __fmul_rn(d.y, t[0])));
  cuv[(int64_t)b * cuv_stride + i] = make_float2(u, v);
}

// uv += resize(cuv, level size, INTER_CUBIC)   (ecc2.cc:2392-2395, 2827-2831)
__global__ void __launch_bounds__(256) k_flow_update(float2 *uv, int64_t uv_stride, int w, int h, const float2 *cuv, int64_t cuv_stride, int cw,
                                                     int ch, FlowCubicAxis tx, FlowCubicAxis ty, const EccFrame *frames) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (x >= w || y >= h) return;
  if (frames && !frames[b].ok) return;
  const float2 *src = cuv + (int64_t)b * cuv_stride;
  const float2 d = cubic_at([&](int sx, int sy) { return __ldg(src + (int64_t)sy * cw + sx); }, cw, ch, __ldg(tx.s + x), __ldg(tx.c + x),
                            __ldg(ty.s + y), __ldg(ty.c + y));
  float2 *p = uv + (int64_t)b * uv_stride + (int64_t)y * w + x;
  const float2 o = *p;
  *p = make_float2(__fadd_rn(o.x, d.x), __fadd_rn(o.y, d.y));
}

// level change: uv_dst = resize(uv_src * ratio, level size, INTER_CUBIC)   (ecc2.cc:2812-2825)
__global__ void __launch_bounds__(256) k_flow_upscale(const float2 *src, int64_t src_stride, int sw, int sh, float rx, float ry, float2 *dst,
                                                      int64_t dst_stride, int w, int h, FlowCubicAxis tx, FlowCubicAxis ty,
                                                      const EccFrame *frames) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (x >= w || y >= h) return;
  if (frames && !frames[b].ok) return;
  const float2 *s = src + (int64_t)b * src_stride;
  dst[(int64_t)b * dst_stride + (int64_t)y * w + x] = cubic_at(
      [&](int sx, int sy) { const float2 f = __ldg(s + (int64_t)sy * sw + sx); return make_float2(__fmul_rn(f.x, rx), __fmul_rn(f.y, ry)); }, sw,
      sh, __ldg(tx.s + x), __ldg(tx.c + x), __ldg(ty.s + y), __ldg(ty.c + y));
}

// initial flow: uv = resize(map - grid, last level size, INTER_CUBIC) * ratio   (ecc2.cc:2797-2800); the map is the analytic
// registration map of the frame (frames[b].map) or an explicit CV_32FC2 map
__global__ void __launch_bounds__(256) k_flow_init(const EccFrame *frames, const float2 *rmap, int sw, int sh, float rx, float ry, float2 *dst,
                                                   int64_t dst_stride, int w, int h, FlowCubicAxis tx, FlowCubicAxis ty) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (x >= w || y >= h) return;
  float2 o;
  if (rmap) {
    o = cubic_at([&](int sx, int sy) { const float2 m = __ldg(rmap + (int64_t)sy * sw + sx);
                                        return make_float2(__fsub_rn(m.x, (float)sx), __fsub_rn(m.y, (float)sy)); },
                 sw, sh, __ldg(tx.s + x), __ldg(tx.c + x), __ldg(ty.s + y), __ldg(ty.c + y));
  } else {
    if (!frames[b].ok) return;
    const MapCoef m = frames[b].map;
    o = cubic_at([&](int sx, int sy) { float u, v; map_xy(m, (float)sx, (float)sy, u, v);
                                        return make_float2(__fsub_rn(u, (float)sx), __fsub_rn(v, (float)sy)); },
                 sw, sh, __ldg(tx.s + x), __ldg(tx.c + x), __ldg(ty.s + y), __ldg(ty.c + y));
  }
  dst[(int64_t)b * dst_stride + (int64_t)y * w + x] = make_float2(__fmul_rn(o.x, rx), __fmul_rn(o.y, ry));
}

__global__ void __launch_bounds__(256) k_flow_to_remap(const float2 *uv, int w, int h, float2 *rmap) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  const float2 f = uv[(int64_t)y * w + x];
  rmap[(int64_t)y * w + x] = make_float2(__fadd_rn(f.x, (float)x), __fadd_rn(f.y, (float)y));
}

// ---------------------------------------------------------------------------------------------------------------------
// host-side tables
// ---------------------------------------------------------------------------------------------------------------------
struct AreaTabHost { std::vector<int> start, count, off; std::vector<float> alpha; };

// computeResizeAreaTab (OpenCV imgproc/resize.cpp) for one axis; an integer scale takes ResizeAreaFast's plain 1 / scale
void area_tab(int ssize, int dsize, AreaTabHost &t) {
  const double scale = 1.0 / ((double)dsize / ssize);
  t.start.assign(dsize, 0); t.count.assign(dsize, 0); t.off.assign(dsize, 0); t.alpha.clear();
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    t.off[dx] = (int)t.alpha.size();
    int first = -1;
    if (sx1 - fsx1 > 1e-3) { first = sx1 - 1; t.alpha.push_back((float)((sx1 - fsx1) / cell)); }
    for (int sx = sx1; sx < sx2; ++sx) { if (first < 0) first = sx; t.alpha.push_back((float)(1.0 / cell)); }
    if (fsx2 - sx2 > 1e-3) { if (first < 0) first = sx2; t.alpha.push_back((float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell)); }
    if (first < 0) { first = std::min(sx1, ssize - 1); t.alpha.push_back(1.f); }
    t.start[dx] = first; t.count[dx] = (int)t.alpha.size() - t.off[dx];
  }
}

struct CubicTabHost { std::vector<int> s; std::vector<float> c; };

// the per-axis part of cv::resize(INTER_CUBIC): fx = (dx + 0.5) * scale - 0.5, interpolateCubic(fx - floor(fx)) in float
void cubic_tab(int ssize, int dsize, CubicTabHost &t) {
  const double scale = 1.0 / ((double)dsize / ssize);
  t.s.resize(dsize); t.c.resize((size_t)dsize * 4);
  for (int dx = 0; dx < dsize; ++dx) {
    float fx = (float)((dx + 0.5) * scale - 0.5);
    const int sx = (int)std::floor(fx);
    fx -= sx;
    const float A = -0.75f;
    float *c = &t.c[(size_t)dx * 4];
    c[0] = ((A * (fx + 1) - 5 * A) * (fx + 1) + 8 * A) * (fx + 1) - 4 * A;
    c[1] = ((A + 2) * fx - (A + 3)) * fx * fx + 1;
    c[2] = ((A + 2) * (1 - fx) - (A + 3)) * (1 - fx) * (1 - fx) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
    t.s[dx] = sx;
  }
}

struct Blob {
  std::vector<unsigned char> bytes;
  size_t add(const void *p, size_t n) {
    const size_t at = (bytes.size() + 15) & ~(size_t)15;
    bytes.resize(at + n);
    memcpy(bytes.data() + at, p, n);
    return at;
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// EccFlow
// ---------------------------------------------------------------------------------------------------------------------
EccFlowHolder::~EccFlowHolder() { delete p; }

int EccFlow::init(const ssk_eccflow_options &o, cudaStream_t s) {
  opts = o; stream = s; have_reference = false;
  SSK_REQUIRE(o.support_scale >= 0 && o.support_scale <= 12, "eccflow: support_scale 0..12");
  SSK_REQUIRE(o.max_iterations >= 0, "eccflow: max_iterations >= 0");
  SSK_REQUIRE(o.scale_factor > 0 && o.scale_factor < 1, "eccflow: scale_factor in (0, 1)");
  SSK_REQUIRE(o.downscale_method >= 0 && o.downscale_method <= 2, "eccflow: unknown downscale method");
  return SSK_OK;
}

int EccFlow::build_tables() {
  Blob blob;
  struct AxOff { size_t start, count, off, alpha; };
  struct CxOff { size_t s, c; };
  auto put_area = [&](int ssize, int dsize) {
    AreaTabHost t; area_tab(ssize, dsize, t);
    AxOff o;
    o.start = blob.add(t.start.data(), t.start.size() * 4); o.count = blob.add(t.count.data(), t.count.size() * 4);
    o.off = blob.add(t.off.data(), t.off.size() * 4); o.alpha = blob.add(t.alpha.data(), t.alpha.size() * 4);
    return o;
  };
  auto put_cubic = [&](int ssize, int dsize) {
    CubicTabHost t; cubic_tab(ssize, dsize, t);
    CxOff o;
    o.s = blob.add(t.s.data(), t.s.size() * 4); o.c = blob.add(t.c.data(), t.c.size() * 4);
    return o;
  };
  std::vector<AxOff> ax(nlevels), ay(nlevels);
  std::vector<CxOff> ux(nlevels), uy(nlevels), nx(nlevels), ny(nlevels);
  for (int l = 0; l < nlevels; ++l) {
    ax[l] = put_area(lw[l], cw[l]); ay[l] = put_area(lh[l], ch[l]);
    ux[l] = put_cubic(cw[l], lw[l]); uy[l] = put_cubic(ch[l], lh[l]);
    if (l + 1 < nlevels) { nx[l] = put_cubic(lw[l + 1], lw[l]); ny[l] = put_cubic(lh[l + 1], lh[l]); }
  }
  const CxOff i_x = put_cubic(lw[0], lw[nlevels - 1]), i_y = put_cubic(lh[0], lh[nlevels - 1]);
  if (int e = tabs.ensure(blob.bytes.size())) return e;
  SSK_CUDA(cudaMemcpyAsync(tabs.p, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice, stream));
  SSK_CUDA(cudaStreamSynchronize(stream));   // the blob is a local
  const char *base = tabs.as<char>();
  auto area = [&](const AxOff &o) {
    FlowAreaAxis a;
    a.start = (const int *)(base + o.start); a.count = (const int *)(base + o.count); a.off = (const int *)(base + o.off);
    a.alpha = (const float *)(base + o.alpha);
    return a;
  };
  auto cubic = [&](const CxOff &o) { FlowCubicAxis c; c.s = (const int *)(base + o.s); c.c = (const float4 *)(base + o.c); return c; };
  for (int l = 0; l < nlevels; ++l) {
    lt[l].ax = area(ax[l]); lt[l].ay = area(ay[l]);
    lt[l].ux = cubic(ux[l]); lt[l].uy = cubic(uy[l]);
    if (l + 1 < nlevels) { lt[l].nx = cubic(nx[l]); lt[l].ny = cubic(ny[l]); }
  }
  ix = cubic(i_x); iy = cubic(i_y);
  return SSK_OK;
}

// downscale() of one level from its source level for `batch` images laid out with the pyramid stride
static int flow_downscale(int method, const float *src_base, float *dst_base, const float *const *src_ptrs, float *const *dst_ptrs, int sw, int sh,
                          int dw, int dh, int batch, cudaStream_t s) {
  if (method == SSK_ECCFLOW_DOWNSCALE_PYRAMID) {
    PyrDownArgs pd = {};
    pd.src.data = src_base; pd.src.step = (int64_t)sw * 4; pd.src.rows = sh; pd.src.cols = sw; pd.src.depth = SSK_32F; pd.src.cn = 1;
    pd.src.scale = 1.f; pd.src_ptrs = reinterpret_cast<const void *const *>(src_ptrs);
    pd.dst = dst_base; pd.dst_ptrs = dst_ptrs; pd.dst_rows = dh; pd.dst_cols = dw; pd.batch = batch; pd.post_scale = 1.f;
    return launch_pyrdown(pd, s);
  }
  ResizeAreaArgs ra = {};
  ra.src.data = src_base; ra.src.step = (int64_t)sw * 4; ra.src.rows = sh; ra.src.cols = sw; ra.src.depth = SSK_32F; ra.src.cn = 1;
  ra.src.scale = 1.f; ra.src_ptrs = reinterpret_cast<const void *const *>(src_ptrs);
  ra.dst = dst_base; ra.dst_ptrs = dst_ptrs; ra.dst_rows = dh; ra.dst_cols = dw; ra.batch = batch;
  ra.inv_scale_x = (double)dw / sw; ra.inv_scale_y = (double)dh / sh;
  return launch_resize_area(ra, s);
}

int EccFlow::set_reference(const float *d_img, int rows, int cols, const uint8_t *d_mask) {
  have_reference = false;
  // level sizes: ecc2.cc:2520-2600
  const int min_image_size = std::max(4, opts.min_image_size);
  const bool big_aspect = std::max(cols, rows) / std::min(cols, rows) >= 2;
  nlevels = 1; lw[0] = cols; lh[0] = rows; lsrc[0] = -1;
  for (int lvl = 1;; ++lvl) {
    if (opts.max_pyramid_level >= 0 && lvl - 1 >= opts.max_pyramid_level) break;
    if (lvl >= kMaxFlowLevels) break;
    const int pw = lw[lvl - 1], ph = lh[lvl - 1];
    int nw, nh, src;
    if (opts.downscale_method != SSK_ECCFLOW_DOWNSCALE_PYRAMID) {
      nw = std::max(opts.min_image_size, (int)((pw + 1) * opts.scale_factor));
      nh = std::max(opts.min_image_size, (int)((ph + 1) * opts.scale_factor));
      if ((nw == pw && nh == ph) || std::max(nw, nh) <= min_image_size) break;
      if (opts.downscale_method == SSK_ECCFLOW_DOWNSCALE_FULL_RESIZE) src = 0;
      else src = (big_aspect && std::min(nw, nh) <= min_image_size + 1) ? 0 : lvl - 1;
    } else {
      nw = std::max(opts.min_image_size, (pw + 1) / 2);
      nh = std::max(opts.min_image_size, (ph + 1) / 2);
      if ((nw == pw && nh == ph) || std::min(nw, nh) <= min_image_size) break;
      src = lvl - 1;
    }
    SSK_REQUIRE(nw >= 1 && nh >= 1, "eccflow: pyramid level collapsed to an empty size");
    lw[lvl] = nw; lh[lvl] = nh; lsrc[lvl] = src;
    nlevels = lvl + 1;
  }
  pyr_px = coarse_px = 0;
  for (int l = 0; l < nlevels; ++l) {
    int w = lw[l], h = lh[l];
    for (int i = 0; i < opts.support_scale; ++i) { w = (w + 1) / 2; h = (h + 1) / 2; }
    cw[l] = w; ch[l] = h;
    loff[l] = pyr_px; coff[l] = coarse_px;
    pyr_px += ((int64_t)lw[l] * lh[l] + 63) & ~(int64_t)63;
    coarse_px += ((int64_t)cw[l] * ch[l] + 15) & ~(int64_t)15;
  }
  SSK_REQUIRE((size_t)lw[0] * 3 * 4 <= 96 * 1024, "eccflow: image wider than 8192 pixels");
  if ((size_t)lw[0] * 3 * 4 > 48 * 1024 && !reduce_smem_optin) {
    SSK_CUDA(cudaFuncSetAttribute(k_flow_reduce<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    SSK_CUDA(cudaFuncSetAttribute(k_flow_reduce<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    SSK_CUDA(cudaFuncSetAttribute(k_flow_reduce<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    reduce_smem_optin = 1;
  }
  if (int e = build_tables()) return e;
  if (int e = ref_pyr.ensure(pyr_px * 4)) return e;
  if (int e = ref_ix.ensure(pyr_px * 4)) return e;
  if (int e = ref_iy.ensure(pyr_px * 4)) return e;
  if (int e = ref_D.ensure(coarse_px * 16)) return e;
  if (int e = raw.ensure((size_t)cw[0] * ch[0] * 3 * 4)) return e;
  float *rp = ref_pyr.as<float>();
  SSK_CUDA(cudaMemcpyAsync(rp, d_img, (size_t)rows * cols * 4, cudaMemcpyDeviceToDevice, stream));
  have_ref_mask = d_mask != nullptr;
  if (have_ref_mask) {
    if (int e = ref_mask.ensure(pyr_px)) return e;
    SSK_CUDA(cudaMemcpyAsync(ref_mask.p, d_mask, (size_t)rows * cols, cudaMemcpyDeviceToDevice, stream));
  }
  const double noise_level = opts.noise_level >= 0 ? opts.noise_level : 1e-3;
  for (int l = 0; l < nlevels; ++l) {
    if (l > 0) {
      const int sl = lsrc[l];
      if (int e = flow_downscale(opts.downscale_method, rp + loff[sl], rp + loff[l], nullptr, nullptr, lw[sl], lh[sl], lw[l], lh[l], 1, stream)) return e;
      if (have_ref_mask)
        if (int e = launch_resize_nearest_u8(ref_mask.as<uint8_t>() + loff[sl], lh[sl], lw[sl], ref_mask.as<uint8_t>() + loff[l], lh[l], lw[l], stream)) return e;
    }
    // ecc_differentiate (ecc2.cc:142-169), no mask
    SepFilterArgs g = {};
    g.rows = lh[l]; g.cols = lw[l]; g.batch = 1; g.src = rp + loff[l];
    g.dst = ref_ix.as<float>() + loff[l];
    g.kxn = 5; g.kyn = 3; memcpy(g.kx, kD5, sizeof(kD5)); memcpy(g.ky, kS3, sizeof(kS3));
    if (int e = launch_sepfilter(g, stream)) return e;
    g.dst = ref_iy.as<float>() + loff[l];
    g.kxn = 3; g.kyn = 5; memcpy(g.kx, kS3, sizeof(kS3)); memcpy(g.ky, kD5, sizeof(kD5));
    if (int e = launch_sepfilter(g, stream)) return e;
    // avgp x 3 + D
    FlowReduceArgs a = {};
    a.w = lw[l]; a.h = lh[l]; a.cw = cw[l]; a.ch = ch[l];
    a.ix = ref_ix.as<float>() + loff[l]; a.iy = ref_iy.as<float>() + loff[l];
    a.out = raw.as<float>(); a.out_stride = 0; a.ax = lt[l].ax; a.ay = lt[l].ay;
    k_flow_reduce<1, 0><<<dim3(ch[l], 1), 256, (size_t)lw[l] * 3 * 4, stream>>>(a);
    SSK_LAUNCH_CHECK();
    // "this regularization term estimation looks crazy" (ecc2.cc:2613): float(pow(1e-5 * noise / 2^level, 4))
    const float reg = noise_level > 0 ? (float)std::pow(1e-5 * noise_level / (double)(1ll << std::min(l, 62)), 4) : 0.f;
    k_flow_D<<<div_up(cw[l] * ch[l], 256), 256, 0, stream>>>(raw.as<float>(), cw[l], ch[l], reg, (float)opts.update_multiplier,
                                                             ref_D.as<float4>() + coff[l]);
    SSK_LAUNCH_CHECK();
  }
  have_reference = true;
  capacity = 0;   // buffers depend on the geometry
  return SSK_OK;
}

int EccFlow::reserve(int batch) {
  SSK_REQUIRE(have_reference, "eccflow: set_reference_image() must be called first");
  if (batch <= capacity) return SSK_OK;
  const int64_t n0 = (int64_t)lw[0] * lh[0];
  if (int e = cur_pyr.ensure((size_t)batch * pyr_px * 4)) return e;
  if (int e = uv_a.ensure((size_t)batch * n0 * 8)) return e;
  if (int e = uv_b.ensure((size_t)batch * n0 * 8)) return e;
  if (int e = raw.ensure(std::max((size_t)batch * cw[0] * ch[0] * 2 * 4, (size_t)cw[0] * ch[0] * 3 * 4))) return e;
  if (int e = cuv.ensure((size_t)batch * cw[0] * ch[0] * 8)) return e;
  // pointer tables: [level][slot]
  std::vector<float *> ptrs((size_t)nlevels * batch);
  for (int l = 0; l < nlevels; ++l)
    for (int b = 0; b < batch; ++b) ptrs[(size_t)l * batch + b] = cur_pyr.as<float>() + (int64_t)b * pyr_px + loff[l];
  if (int e = d_lvl_ptrs.ensure(ptrs.size() * sizeof(float *))) return e;
  SSK_CUDA(cudaMemcpyAsync(d_lvl_ptrs.p, ptrs.data(), ptrs.size() * sizeof(float *), cudaMemcpyHostToDevice, stream));
  if (int e = d_cur_ptrs.ensure((size_t)batch * sizeof(float *))) return e;
  SSK_CUDA(cudaMemcpyAsync(d_cur_ptrs.p, ptrs.data(), (size_t)batch * sizeof(float *), cudaMemcpyHostToDevice, stream));
  SSK_CUDA(cudaStreamSynchronize(stream));
  capacity = batch;
  return SSK_OK;
}

int EccFlow::build_current(int batch, const uint8_t *d_mask) {
  SSK_REQUIRE(batch >= 1 && batch <= capacity, "eccflow: batch exceeds the reserved capacity");
  SSK_REQUIRE(!d_mask || batch == 1, "eccflow: a current mask needs a single-frame call");
  cur_mask_set = d_mask != nullptr;
  if (cur_mask_set) {
    if (int e = cur_mask.ensure(pyr_px)) return e;
    SSK_CUDA(cudaMemcpyAsync(cur_mask.p, d_mask, (size_t)lw[0] * lh[0], cudaMemcpyDeviceToDevice, stream));
  }
  float *const *tab = d_lvl_ptrs.as<float *>();
  for (int l = 1; l < nlevels; ++l) {
    const int sl = lsrc[l];
    if (int e = flow_downscale(opts.downscale_method, nullptr, nullptr, tab + (size_t)sl * capacity, tab + (size_t)l * capacity, lw[sl], lh[sl],
                               lw[l], lh[l], batch, stream)) return e;
    if (cur_mask_set)
      if (int e = launch_resize_nearest_u8(cur_mask.as<uint8_t>() + loff[sl], lh[sl], lw[sl], cur_mask.as<uint8_t>() + loff[l], lh[l], lw[l], stream)) return e;
  }
  return SSK_OK;
}

int EccFlow::compute(int batch, const EccFrame *d_frames, const float2 *d_rmap0) {
  SSK_REQUIRE(batch >= 1 && batch <= capacity, "eccflow: batch exceeds the reserved capacity");
  SSK_REQUIRE(!d_rmap0 || batch == 1, "eccflow: an explicit initial map needs a single-frame call");
  const int64_t n0 = (int64_t)lw[0] * lh[0];
  const int L = nlevels - 1;
  // the flow of level l lives in uv_a when (L - l) is even ... simpler: ping-pong and copy at the end if needed
  float2 *cur_uv = uv_a.as<float2>(), *other = uv_b.as<float2>();
  if (d_frames || d_rmap0) {
    const float rx = (float)((double)lw[L] / (double)lw[0]), ry = (float)((double)lh[L] / (double)lh[0]);
    dim3 grid(div_up(lw[L], 32), div_up(lh[L], 8), batch);
    k_flow_init<<<grid, 256, 0, stream>>>(d_frames, d_rmap0, lw[0], lh[0], rx, ry, cur_uv, n0, lw[L], lh[L], ix, iy);
    SSK_LAUNCH_CHECK();
  } else {
    SSK_CUDA(cudaMemsetAsync(cur_uv, 0, (size_t)batch * n0 * 8, stream));
  }
  const EccFrame *okf = d_rmap0 ? nullptr : d_frames;
  for (int l = L; l >= 0; --l) {
    if (l < L) {
      const float rx = (float)((double)lw[l] / (double)lw[l + 1]), ry = (float)((double)lh[l] / (double)lh[l + 1]);
      dim3 grid(div_up(lw[l], 32), div_up(lh[l], 8), batch);
      k_flow_upscale<<<grid, 256, 0, stream>>>(cur_uv, n0, lw[l + 1], lh[l + 1], rx, ry, other, n0, lw[l], lh[l], lt[l].nx, lt[l].ny, okf);
      SSK_LAUNCH_CHECK();
      std::swap(cur_uv, other);
    }
    // iterations after the first apply the previous update inside k_flow_reduce (read uv, write the other buffer)
    static const bool fuse_update = !getenv("SSK_FLOW_NO_FUSED_UPDATE");
    for (int j = 0; j < opts.max_iterations; ++j) {
      FlowReduceArgs a = {};
      a.w = lw[l]; a.h = lh[l]; a.cw = cw[l]; a.ch = ch[l];
      a.ref = ref_pyr.as<float>() + loff[l]; a.ix = ref_ix.as<float>() + loff[l]; a.iy = ref_iy.as<float>() + loff[l];
      a.refmask = have_ref_mask ? ref_mask.as<uint8_t>() + loff[l] : nullptr;
      a.cur = cur_pyr.as<float>() + loff[l]; a.cur_stride = pyr_px;
      a.curmask = cur_mask_set ? cur_mask.as<uint8_t>() + loff[l] : nullptr;
      a.uv = cur_uv; a.uv_stride = n0;
      a.out = raw.as<float>(); a.out_stride = (int64_t)cw[l] * ch[l] * 2;
      a.ax = lt[l].ax; a.ay = lt[l].ay; a.frames = okf;
      if (fuse_update && j > 0) {
        a.cuv_prev = cuv.as<float2>(); a.cuv_stride = (int64_t)cw[l] * ch[l]; a.ux = lt[l].ux; a.uy = lt[l].uy; a.uv_out = other;
      }
      if (a.cuv_prev) k_flow_reduce<0, 1><<<dim3(ch[l], batch), 256, (size_t)lw[l] * 2 * 4, stream>>>(a);
      else k_flow_reduce<0, 0><<<dim3(ch[l], batch), 256, (size_t)lw[l] * 2 * 4, stream>>>(a);
      SSK_LAUNCH_CHECK();
      if (a.cuv_prev) std::swap(cur_uv, other);
      k_flow_solve<<<dim3(div_up(cw[l] * ch[l], 256), batch), 256, 0, stream>>>(raw.as<float>(), a.out_stride, ref_D.as<float4>() + coff[l], cw[l],
                                                                              ch[l], cuv.as<float2>(), (int64_t)cw[l] * ch[l], okf);
      SSK_LAUNCH_CHECK();
      if (!fuse_update || j == opts.max_iterations - 1) {
        dim3 grid(div_up(lw[l], 32), div_up(lh[l], 8), batch);
        k_flow_update<<<grid, 256, 0, stream>>>(cur_uv, n0, lw[l], lh[l], cuv.as<float2>(), (int64_t)cw[l] * ch[l], cw[l], ch[l], lt[l].ux, lt[l].uy,
                                                okf);
        SSK_LAUNCH_CHECK();
      }
    }
  }
  if (cur_uv != uv_a.as<float2>())
    SSK_CUDA(cudaMemcpyAsync(uv_a.p, cur_uv, (size_t)batch * n0 * 8, cudaMemcpyDeviceToDevice, stream));
  return SSK_OK;
}

int EccFlow::write_remap(int b, float2 *d_rmap) {
  dim3 grid(div_up(lw[0], 32), div_up(lh[0], 8));
  k_flow_to_remap<<<grid, 256, 0, stream>>>(uv(b), lw[0], lh[0], d_rmap);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

const float *EccFlow::image(int which, int level) const {
  if (level < 0 || level >= nlevels) return nullptr;
  switch (which) {
    case 0: return ref_pyr.as<float>() + loff[level];
    case 1: return cur_pyr.as<float>() + loff[level];
    case 2: return ref_ix.as<float>() + loff[level];
    case 3: return ref_iy.as<float>() + loff[level];
    case 4: return reinterpret_cast<const float *>(ref_D.as<float4>() + coff[level]);
  }
  return nullptr;
}

}  // namespace ssk

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
using namespace ssk;

namespace {

int flow_check_mat(const ssk_mat *m, const char *what) {
  if (!m || !m->data || m->rows <= 0 || m->cols <= 0) { set_error(std::string(what) + ": empty image"); return SSK_ERR_INVALID; }
  const int d = type_depth(m->type), cn = type_cn(m->type);
  if (!depth_bytes(d) || cn < 1 || cn > 4) { set_error(std::string(what) + ": unsupported type"); return SSK_ERR_INVALID; }
  if (m->step < (int64_t)m->cols * cn * depth_bytes(d)) { set_error(std::string(what) + ": step smaller than a row"); return SSK_ERR_INVALID; }
  return SSK_OK;
}

// single-channel image of any depth -> dense CV_32FC1 on the device (convertTo(CV_32F), no scaling: ecc2.cc:2225)
int flow_image_to_device(ssk_eccflow *h, const ssk_mat *m, float *d_dst) {
  Img im;
  const int d = type_depth(m->type);
  im.rows = m->rows; im.cols = m->cols; im.depth = d; im.cn = 1; im.scale = 1.f;
  if (m->mem == SSK_MEM_DEVICE) { im.data = m->data; im.step = m->step; }
  else {
    const size_t rowb = (size_t)m->cols * depth_bytes(d);
    if (int e = h->st_img.ensure(rowb * m->rows)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(h->st_img.p, rowb, m->data, m->step, rowb, m->rows, cudaMemcpyHostToDevice, h->stream));
    im.data = h->st_img.p; im.step = (int64_t)rowb;
  }
  return launch_to_gray(im, nullptr, d_dst, nullptr, 1, h->stream);
}

int flow_mask_to_device(ssk_eccflow *h, const ssk_mat *mask, int rows, int cols, const uint8_t **out) {
  *out = nullptr;
  if (!mask || !mask->data) return SSK_OK;
  if (int e = flow_check_mat(mask, "eccflow mask")) return e;
  SSK_REQUIRE(mask->type == SSK_8UC1 && mask->rows == rows && mask->cols == cols, "eccflow: mask must be CV_8UC1 of the image size");
  if (int e = h->st_mask.ensure((size_t)rows * cols)) return e;
  SSK_CUDA(cudaMemcpy2DAsync(h->st_mask.p, cols, mask->data, mask->step, cols, rows,
                             mask->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  *out = h->st_mask.as<uint8_t>();
  return SSK_OK;
}

int flow_have_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    set_error(std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
    cudaGetLastError();
    return SSK_ERR_CUDA;
  }
  return SSK_OK;
}

}  // namespace

extern "C" {

void ssk_eccflow_options_default(ssk_eccflow_options *o) {
  memset(o, 0, sizeof(*o));
  o->update_multiplier = 1.5; o->scale_factor = 0.5; o->noise_level = -1; o->max_iterations = 1; o->support_scale = 5;
  o->min_image_size = 4; o->max_pyramid_level = -1; o->downscale_method = SSK_ECCFLOW_DOWNSCALE_RECURSIVE_RESIZE;
}

void ssk_eccflow_registration_options_default(ssk_eccflow_options *o) {
  memset(o, 0, sizeof(*o));
  o->update_multiplier = 1.5; o->scale_factor = 0.75; o->noise_level = -1; o->max_iterations = 3; o->support_scale = 4;
  o->min_image_size = -1; o->max_pyramid_level = -1; o->downscale_method = SSK_ECCFLOW_DOWNSCALE_RECURSIVE_RESIZE;
}

int ssk_eccflow_create(const ssk_eccflow_options *opts, ssk_eccflow **out) {
  if (int e = flow_have_device()) return e;
  SSK_REQUIRE(opts && out, "null argument");
  ssk_eccflow *h = new (std::nothrow) ssk_eccflow();
  SSK_REQUIRE(h, "out of memory");
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; set_error("cudaStreamCreate failed"); return SSK_ERR_CUDA; }
  if (int e = h->f.init(*opts, h->stream)) { delete h; return e; }
  *out = h;
  return SSK_OK;
}

int ssk_eccflow_destroy(ssk_eccflow *h) { delete h; return SSK_OK; }

int ssk_eccflow_set_reference_image(ssk_eccflow *h, const ssk_mat *image, const ssk_mat *mask) {
  SSK_REQUIRE(h, "null handle");
  if (int e = flow_check_mat(image, "eccflow reference image")) return e;
  SSK_REQUIRE(type_cn(image->type) == 1, "eccflow: single channel image expected");   // ecc2.cc:2498-2502
  if (int e = h->st_map.ensure((size_t)image->rows * image->cols * 4)) return e;
  if (int e = flow_image_to_device(h, image, h->st_map.as<float>())) return e;
  const uint8_t *dm;
  if (int e = flow_mask_to_device(h, mask, image->rows, image->cols, &dm)) return e;
  if (int e = h->f.set_reference(h->st_map.as<float>(), image->rows, image->cols, dm)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

int ssk_eccflow_compute(ssk_eccflow *h, const ssk_mat *image, const ssk_mat *mask, ssk_mat *rmap, int use_initial_map) {
  SSK_REQUIRE(h, "null handle");
  SSK_REQUIRE(h->f.have_reference, "eccflow: set_reference_image() must be called first");   // ecc2.cc:2678-2681
  if (int e = flow_check_mat(image, "eccflow input image")) return e;
  SSK_REQUIRE(type_cn(image->type) == 1, "eccflow: single channel image expected");
  SSK_REQUIRE(image->rows == h->f.lh[0] && image->cols == h->f.lw[0], "eccflow: input image size differs from the reference");
  if (int e = flow_check_mat(rmap, "eccflow rmap")) return e;
  // ecc2.cc:2787-2805: a non-empty rmap must have the reference size
  SSK_REQUIRE(rmap->type == SSK_32FC2 && rmap->rows == h->f.lh[0] && rmap->cols == h->f.lw[0], "eccflow: rmap must be CV_32FC2 of the reference size");
  if (int e = h->f.reserve(1)) return e;
  if (int e = flow_image_to_device(h, image, h->f.level0(0))) return e;
  const uint8_t *dm;
  if (int e = flow_mask_to_device(h, mask, image->rows, image->cols, &dm)) return e;
  if (int e = h->f.build_current(1, dm)) return e;
  const size_t rowb = (size_t)rmap->cols * 8;
  if (int e = h->st_map.ensure(rowb * rmap->rows)) return e;
  const float2 *d_rmap0 = nullptr;
  if (use_initial_map) {
    SSK_CUDA(cudaMemcpy2DAsync(h->st_map.p, rowb, rmap->data, rmap->step, rowb, rmap->rows,
                               rmap->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    d_rmap0 = h->st_map.as<float2>();
  }
  if (int e = h->f.compute(1, nullptr, d_rmap0)) return e;
  if (int e = h->f.write_remap(0, h->st_map.as<float2>())) return e;
  SSK_CUDA(cudaMemcpy2DAsync(rmap->data, rmap->step, h->st_map.p, rowb, rowb, rmap->rows,
                             rmap->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

int ssk_eccflow_get_uv(ssk_eccflow *h, ssk_mat *uv) {
  SSK_REQUIRE(h && h->f.have_reference && h->f.capacity >= 1, "eccflow: no flow computed yet");
  if (int e = flow_check_mat(uv, "eccflow uv")) return e;
  SSK_REQUIRE(uv->type == SSK_32FC2 && uv->rows == h->f.lh[0] && uv->cols == h->f.lw[0], "eccflow: uv must be CV_32FC2 of the reference size");
  const size_t rowb = (size_t)uv->cols * 8;
  SSK_CUDA(cudaMemcpy2DAsync(uv->data, uv->step, h->f.uv(0), rowb, rowb, uv->rows,
                             uv->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

int ssk_eccflow_num_levels(const ssk_eccflow *h) { return h && h->f.have_reference ? h->f.nlevels : 0; }

int ssk_eccflow_level_size(const ssk_eccflow *h, int level, int *cols, int *rows, int *grid_cols, int *grid_rows) {
  SSK_REQUIRE(h && h->f.have_reference && level >= 0 && level < h->f.nlevels, "eccflow: no such level");
  if (cols) *cols = h->f.lw[level];
  if (rows) *rows = h->f.lh[level];
  if (grid_cols) *grid_cols = h->f.cw[level];
  if (grid_rows) *grid_rows = h->f.ch[level];
  return SSK_OK;
}

int ssk_eccflow_get_image(ssk_eccflow *h, int which, int level, ssk_mat *dst) {
  SSK_REQUIRE(h && h->f.have_reference && level >= 0 && level < h->f.nlevels, "eccflow: no such level");
  SSK_REQUIRE(which >= 0 && which <= 4, "eccflow: which = 0..4");
  SSK_REQUIRE(which != 1 || h->f.capacity >= 1, "eccflow: no current image yet");
  if (int e = flow_check_mat(dst, "eccflow get_image")) return e;
  const int w = which == 4 ? h->f.cw[level] : h->f.lw[level], hh = which == 4 ? h->f.ch[level] : h->f.lh[level];
  SSK_REQUIRE(dst->type == (which == 4 ? SSK_32FC4 : SSK_32FC1) && dst->rows == hh && dst->cols == w, "eccflow: dst must match the level geometry");
  const size_t rowb = (size_t)w * (which == 4 ? 16 : 4);
  SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, h->f.image(which, level), rowb, rowb, hh,
                             dst->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

}  // extern "C"
