// K5, per-pixel-map form: fused warp + eroded validity mask + weight-map warp + weighted running mean for a batch of frames
// whose registration map was refined by c_eccflow (the map of frame j is flow_j + grid, a gathered CV_32FC2 field instead of
// an analytic transform).
//
// Reference semantics: c_frame_registration::custom_remap(_current_remap, ...) -> base_remap (c_frame_registration.cc:1265-1417)
// with the map c_eccflow::compute left in _current_remap (c_frame_registration.cc:900-917), then compute_weights /
// multiply_weights / c_weigthed_average::add as on the analytic path (c_image_stacking_pipeline.cc:1644-1779).
//
// The same kernel serves frame_upscale_after_align (c_image_stacking_pipeline.cc:1633-1660): the accumulator is 1.5 / 2 / 3 times
// the frame size and the map of an output pixel is upscale_remap(current_remap) - cv::resize(INTER_LINEAR) or cv::pyrUp of the
// map - evaluated on the fly from the flow field or from the analytic transform (ssk_upscale.cuh); the up-scaled map is
// never written to memory.
//
// Structure: one CTA per 32 x 32 accumulator tile, 256 threads, 4 pixels per thread; mean and weight sum stay in registers
// for all frames of the batch (read and written once per batch).  Per frame the CTA
//   1. reads the flow of the tile plus its 2-px erosion halo (36 x 36 float2, the only per-pixel map traffic), forms the map
//      coordinate and the pre-erosion validity flag remap(all-255, interp, CONSTANT 0) >= 255 of every halo pixel into one of
//      two shared flag buffers (one CTA barrier per frame),
//   2. erodes the flags 5 x 5 (positions outside the image do not erode: border value 255) and, for valid pixels, gathers the
//      weight map and the frame at the map coordinate with cv::remap's arithmetic (sample_any) and updates the running mean.
#include "ssk_fused_impl.cuh"
#include "ssk_upscale.cuh"

namespace ssk {

namespace {

constexpr int HW = TW + 4, HH = TH + 4;      // tile + 2-px halo

// map coordinate of output pixel (gx, gy): current_remap, or its up-scaling
__device__ __forceinline__ float2 map_at(const WarpAccArgs &a, const UpscaleGeom &up, const MapCoef &m, const float2 *flow, int gx, int gy) {
  auto src_uv = [&](int x, int y) -> float2 {
    if (flow) {
      const float2 d = __ldg(flow + (int64_t)y * a.src_cols + x);
      return make_float2(__fadd_rn(d.x, (float)x), __fadd_rn(d.y, (float)y));      // ecc_flow_to_remap
    }
    float u, v;
    map_xy(m, (float)x, (float)y, u, v);
    return make_float2(u, v);
  };
  if (up.option == SSK_UPSCALE_NONE) return src_uv(gx, gy);
  return make_float2(upscale_sample(up, [&](int x, int y) { return src_uv(x, y).x; }, gx, gy),
                     upscale_sample(up, [&](int x, int y) { return src_uv(x, y).y; }, gx, gy));
}

// FAST: 0 = any depth / channel count / interpolation through sample_any (real calls); 1 / 2 = CV_32FC1 frames with INTER_LINEAR /
// INTER_CUBIC: the samplers of ssk_common.cuh inlined (same arithmetic, the interior 4 x 4 footprint without border logic)
template <int FAST>
__global__ void __launch_bounds__(256) k_fused_flow(const __grid_constant__ WarpAccArgs a, const __grid_constant__ Tables tab, int ntx,
                                                    const __grid_constant__ UpscaleGeom up) {
  __shared__ unsigned char s_flag[2][HH][HW + 4];
  __shared__ float2 s_uv[TH][TW];            // map coordinates of the tile's own pixels
  const int bx0 = (blockIdx.x % ntx) * TW, by0 = (blockIdx.x / ntx) * TH;
  const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
  const int x = bx0 + lx;
  float A[4][4], W[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = by0 + ly0 + 8 * k;
    W[k] = 0.f;
    for (int c = 0; c < 4; ++c) A[k][c] = 0.f;
    if (x < a.cols && y < a.rows) {
      const int64_t p = (int64_t)y * a.cols + x;
      W[k] = a.wacc[p];
      for (int c = 0; c < a.cn; ++c) A[k][c] = a.acc[p * a.cn + c];
    }
  }
  int buf = 0;
#pragma unroll 1
  for (int j = 0; j < a.njobs; ++j) {
    const FrameJob &job = a.jobs[j];
    if (!job.ok) continue;                                       // block-uniform
    const float2 *flow = a.flow ? a.flow + (int64_t)j * a.flow_stride : nullptr;
    const MapCoef m = job.map;
    // ---- 1. flags of tile + halo
    for (int i = threadIdx.x; i < HH * HW; i += 256) {
      const int hy = i / HW, hx = i - hy * HW;
      const int gx = bx0 - 2 + hx, gy = by0 - 2 + hy;
      unsigned char f = 1;                                       // outside the image: does not erode
      if ((unsigned)gx < (unsigned)a.cols && (unsigned)gy < (unsigned)a.rows) {
        const float2 uvh = map_at(a, up, m, flow, gx, gy);
        const float u = uvh.x, v = uvh.y;
        f = valid255_t(a.interp, u, v, a.src_cols, a.src_rows, tab) ? 1 : 0;
        if (hx >= 2 && hx < 2 + TW && hy >= 2 && hy < 2 + TH) s_uv[hy - 2][hx - 2] = make_float2(u, v);
      }
      s_flag[buf][hy][hx] = f;
    }
    __syncthreads();
    // ---- 2. erode + sample + accumulate
    const bool weighted = a.use_weights && job.weights != nullptr;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int ly = ly0 + 8 * k, y = by0 + ly;
      if (x >= a.cols || y >= a.rows) continue;
      unsigned ok = 1;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) ok &= s_flag[buf][ly + dy][lx + dx];
      if (!ok) continue;
      const float2 uv = s_uv[ly][lx];
      Img im;
      im.rows = a.src_rows; im.cols = a.src_cols;
      float wk = 1.f;
      if (weighted) {
        im.data = job.weights; im.step = a.w_step; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
        if (FAST == 2) wk = sample_cubic<SSK_32F>(im, 0, uv.x, uv.y, SSK_BORDER_CONSTANT, 0.f, tab.cubic);
        else if (FAST == 1) wk = sample_linear<SSK_32F>(im, 0, uv.x, uv.y, SSK_BORDER_CONSTANT, 0.f);
        else wk = sample_any(im, 0, uv.x, uv.y, a.interp, SSK_BORDER_CONSTANT, 0.f, tab.cubic, tab.lanczos);
        if (!(wk > 0.f)) continue;                               // c_frame_accumulation.cc:114
      }
      im.data = job.frame; im.step = a.src_step; im.depth = a.depth; im.cn = a.cn; im.scale = a.scale;
      const float Wn = W[k] + wk;
      const float factor = weighted ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
      W[k] = Wn;
      if (FAST) {
        const float I = FAST == 2 ? sample_cubic<SSK_32F>(im, 0, uv.x, uv.y, a.border, a.bval[0], tab.cubic)
                                  : sample_linear<SSK_32F>(im, 0, uv.x, uv.y, a.border, a.bval[0]);
        A[k][0] = fmaf(I - A[k][0], factor, A[k][0]);
      } else {
        for (int c = 0; c < a.cn; ++c) {
          const float I = sample_any(im, c, uv.x, uv.y, a.interp, a.border, a.bval[c], tab.cubic, tab.lanczos);
          A[k][c] = fmaf(I - A[k][c], factor, A[k][c]);
        }
      }
    }
    buf ^= 1;
    // s_uv is rewritten by step 1 of the next frame: every thread must have read its entries
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y = by0 + ly0 + 8 * k;
    if (x < a.cols && y < a.rows) {
      const int64_t p = (int64_t)y * a.cols + x;
      a.wacc[p] = W[k];
      for (int c = 0; c < a.cn; ++c) a.acc[p * a.cn + c] = A[k][c];
    }
  }
}

}  // namespace

int launch_warp_accumulate_flow(const WarpAccArgs &a, const Tables &tab, int upscale_option, cudaStream_t s) {
  SSK_REQUIRE(a.flow || upscale_option != SSK_UPSCALE_NONE, "internal: per-pixel-map form without a flow field or an up-scaling");
  const UpscaleGeom up = make_upscale_geom(upscale_option, a.src_cols, a.src_rows);
  SSK_REQUIRE(a.rows == up.dh && a.cols == up.dw, "internal: accumulator size differs from the (up-scaled) map size");
  const int ntx = div_up(a.cols, TW), nty = div_up(a.rows, TH);
  const bool f32c1 = a.depth == SSK_32F && a.cn == 1;
  if (f32c1 && a.interp == SSK_INTER_CUBIC) k_fused_flow<2><<<ntx * nty, 256, 0, s>>>(a, tab, ntx, up);
  else if (f32c1 && a.interp == SSK_INTER_LINEAR) k_fused_flow<1><<<ntx * nty, 256, 0, s>>>(a, tab, ntx, up);
  else k_fused_flow<0><<<ntx * nty, 256, 0, s>>>(a, tab, ntx, up);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk
