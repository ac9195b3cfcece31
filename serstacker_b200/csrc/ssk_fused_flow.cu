// K5, per-pixel-map form: fused warp + eroded validity mask + weight-map warp + weighted running mean for a batch of frames
// whose registration map was refined by c_eccflow (the map of frame j is flow_j + grid, a gathered CV_32FC2 field instead of
// an analytic transform).
//
// Reference semantics: c_frame_registration::custom_remap(_current_remap, ...) -> base_remap (c_frame_registration.cc:1265-1417)
// with the map c_eccflow::compute left in _current_remap (c_frame_registration.cc:900-917), then compute_weights /
// multiply_weights / c_weigthed_average::add as on the analytic path (c_image_stacking_pipeline.cc:1644-1779).
//
// Structure: one CTA per 32 x 32 accumulator tile, 256 threads, 4 pixels per thread; mean and weight sum stay in registers
// for all frames of the batch (read and written once per batch).  Per frame the CTA
//   1. reads the flow of the tile plus its 2-px erosion halo (36 x 36 float2, the only per-pixel map traffic), forms the map
//      coordinate and the pre-erosion validity flag remap(all-255, interp, CONSTANT 0) >= 255 of every halo pixel into one of
//      two shared flag buffers (one CTA barrier per frame),
//   2. erodes the flags 5 x 5 (positions outside the image do not erode: border value 255) and, for valid pixels, gathers the
//      weight map and the frame at the map coordinate with cv::remap's arithmetic (sample_any) and updates the running mean.
#include "ssk_fused_impl.cuh"

namespace ssk {

namespace {

constexpr int HW = TW + 4, HH = TH + 4;      // tile + 2-px halo

__global__ void __launch_bounds__(256) k_fused_flow(const __grid_constant__ WarpAccArgs a, const __grid_constant__ Tables tab, int ntx) {
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
    const float2 *flow = a.flow + (int64_t)j * a.flow_stride;
    // ---- 1. flags of tile + halo
    for (int i = threadIdx.x; i < HH * HW; i += 256) {
      const int hy = i / HW, hx = i - hy * HW;
      const int gx = bx0 - 2 + hx, gy = by0 - 2 + hy;
      unsigned char f = 1;                                       // outside the image: does not erode
      if ((unsigned)gx < (unsigned)a.cols && (unsigned)gy < (unsigned)a.rows) {
        const float2 d = __ldg(flow + (int64_t)gy * a.cols + gx);
        const float u = __fadd_rn(d.x, (float)gx), v = __fadd_rn(d.y, (float)gy);   // ecc_flow_to_remap
        f = valid255(a.interp, u, v, a.src_cols, a.src_rows, tab.cubic_itab) ? 1 : 0;
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
        wk = sample_any(im, 0, uv.x, uv.y, a.interp, SSK_BORDER_CONSTANT, 0.f, tab.cubic);
        if (!(wk > 0.f)) continue;                               // c_frame_accumulation.cc:114
      }
      im.data = job.frame; im.step = a.src_step; im.depth = a.depth; im.cn = a.cn; im.scale = a.scale;
      const float Wn = W[k] + wk;
      const float factor = weighted ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
      W[k] = Wn;
      for (int c = 0; c < a.cn; ++c) {
        const float I = sample_any(im, c, uv.x, uv.y, a.interp, a.border, a.bval[c], tab.cubic);
        A[k][c] = fmaf(I - A[k][c], factor, A[k][c]);
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

int launch_warp_accumulate_flow(const WarpAccArgs &a, const Tables &tab, cudaStream_t s) {
  SSK_REQUIRE(a.flow, "internal: flow form without a flow field");
  SSK_REQUIRE(a.rows == a.src_rows && a.cols == a.src_cols, "eccflow maps have the reference frame size");
  const int ntx = div_up(a.cols, TW), nty = div_up(a.rows, TH);
  k_fused_flow<<<ntx * nty, 256, 0, s>>>(a, tab, ntx);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk
