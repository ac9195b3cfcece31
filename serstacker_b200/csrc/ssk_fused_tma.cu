// K5, TMA form: fused sub-pixel warp + eroded validity mask + weight-map warp + weighted running mean for a batch of
// CV_32F single-channel frames, ONE launch per batch over all tiles of the accumulator (border tiles first).
//
// Reference semantics: as ssk_fused.cu (c_frame_registration::base_remap, c_frame_registration.cc:1265-1386;
// c_image_stacking_pipeline.cc:1653-1660, 1704-1714; _weighted_average_update, c_frame_accumulation.cc:20-129).
//
// Structure
//   * 32 x 32 accumulator tile per CTA (4 warps x 8-row strips), resident in shared memory for all frames of the batch.
//   * The source window of frame k (40 x 40 floats of the frame and of its weight map) is fetched by the tensor copy
//     engine (cp.async.bulk.tensor.2d) into one of two buffers; completion is counted on a transaction mbarrier per
//     buffer.  Out-of-bounds elements arrive as zeros: exactly cv::remap(weights, BORDER_CONSTANT 0), and for the frame
//     every tap outside the frame meets a zero bicubic weight or an invalid (masked) pixel - except the 1-px ring just
//     outside the frame, which can carry the +-23/32768 outer tap of a pixel whose fraction is 1/32 or 31/32; that
//     ring is patched in shared memory with its cv::borderInterpolate value before use (warp_patch).
//   * The four warps of a CTA are decoupled: a warp that has finished a buffer bumps a shared counter, and the last
//     one to arrive re-arms the buffer's barrier and issues the copy of the frame two steps ahead.  There is no CTA
//     barrier inside the frame loop.
//   * base_remap's mask (erode5x5 of the remap validity) is evaluated per warp for its own strip with ballots and
//     kept as one bit per row in a register; tiles whose footprint (with the erosion halo) is inside the frame for a
//     given frame skip it.
#include "ssk_fused_impl.cuh"

namespace ssk {

namespace {

constexpr int NW = 4;                          // warps per CTA
constexpr int GR = TH / NW;                    // rows per warp strip
constexpr int WD = StageGeom<SSK_32F>::WD;     // staged window width (floats)
static_assert(WD == WWD, "frame and weight windows share one geometry on the TMA path");
constexpr unsigned WIN_BYTES = GSH * WD * 4;

// staged: -1 frame dropped by the registration, 0 generic per-pixel path, 1 staged + mask flags, 2 staged, mask all ones
__device__ __noinline__ StagePlan plan_tma(const MapCoef &m, int bx0, int by0, const WarpAccArgs &a) {
  StagePlan p; p.staged = 0; p.sx0 = p.sy0 = p.sxw = 0;
  if (!is_affine_like(m.type) || a.border == SSK_BORDER_WRAP) return p;
  const int cx1 = min(bx0 + TW - 1, a.cols - 1), cy1 = min(by0 + TH - 1, a.rows - 1);
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    float u, v;
    map_xy(m, (float)((k & 1) ? cx1 : bx0), (float)((k & 2) ? cy1 : by0), u, v);
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  if (!(umin > -30000.f && vmin > -30000.f && umax < 30000.f && vmax < 30000.f)) return p;   // plans are packed as shorts
  const int x_lo = (int)floorf(umin) - 1, x_hi = (int)floorf(umax) + 3;
  const int y_lo = (int)floorf(vmin) - 1, y_hi = (int)floorf(vmax) + 3;
  // the copy engine wants the box origin 16-byte aligned along the row (measured: tools/probes/tma_probe.cu traps on a
  // start column that is not a multiple of 4 floats; negative and fully out-of-range origins are fine and zero-filled)
  const int sx0 = x_lo & ~3;
  if (!(x_hi - sx0 < WD && y_hi - y_lo < GSH)) return p;
  p.sx0 = p.sxw = sx0; p.sy0 = y_lo; p.staged = 1;
  // staged = 2: every pixel of the tile and of its 2-px erosion halo (clipped to the image) maps into the tap-safe
  // interior of the frame: the eroded validity mask is all ones for this frame and no tap leaves the frame
  const int hx0 = max(bx0 - 2, 0), hy0 = max(by0 - 2, 0), hx1 = min(bx0 + TW + 1, a.cols - 1), hy1 = min(by0 + TH + 1, a.rows - 1);
  float hu0 = 3.4e38f, hu1 = -3.4e38f, hv0 = 3.4e38f, hv1 = -3.4e38f;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    float u, v;
    map_xy(m, (float)((k & 1) ? hx1 : hx0), (float)((k & 2) ? hy1 : hy0), u, v);
    hu0 = fminf(hu0, u); hu1 = fmaxf(hu1, u); hv0 = fminf(hv0, v); hv1 = fmaxf(hv1, v);
  }
  if (hu0 >= 3.f && hv0 >= 3.f && hu1 <= (float)(a.src_cols - 4) && hv1 <= (float)(a.src_rows - 4)) p.staged = 2;
  return p;
}

// base_remap's mask for the GR rows of a warp's strip: bit k of the result = mask(bx0 + lane, y0 + k).
// Pre-erosion flags of the rows y0 - 2 .. y0 + GR + 1 and the columns bx0 - 2 .. bx0 + TW + 1 are packed by ballots
// (positions outside the image do not erode: border value 255), eroded along x by shifts and along y by a sliding AND.
template <int INTERP, int MT>
__device__ __forceinline__ unsigned warp_okbits(const MapCoef &m, int bx0, int y0, int lane, const WarpAccArgs &a, const short *itab) {
  auto flag = [&](int gx, int gy) -> bool {
    if (gx < 0 || gy < 0 || gx >= a.cols || gy >= a.rows) return true;
    float u, v;
    const ColMap<MT> cmf(m, (float)gx);
    cmf((float)gy, u, v);
    return valid255(INTERP, u, v, a.src_cols, a.src_rows, itab);
  };
  // the four trailing columns (30 .. 33 of the tile) of all GR + 4 rows: (row, column) pairs across the lanes
  static_assert(GR + 4 <= 16, "two ballots cover the trailing columns");
  const int er = lane >> 2, ex = bx0 + 30 + (lane & 3);
  const unsigned ex0 = __ballot_sync(0xffffffffu, flag(ex, y0 - 2 + er));
  const unsigned ex1 = __ballot_sync(0xffffffffu, er + 8 < GR + 4 ? flag(ex, y0 - 2 + er + 8) : true);
  unsigned h0 = 0, h1 = 0, h2 = 0, h3 = 0, ok = 0;
#pragma unroll 1      // kept rolled: the frame loop's instruction footprint has to stay inside the 32 KB L1.5 I-cache
  for (int r = 0; r < GR + 4; ++r) {
    const unsigned lo = __ballot_sync(0xffffffffu, flag(bx0 - 2 + lane, y0 - 2 + r));
    const unsigned hi = ((r < 8 ? ex0 >> (4 * r) : ex1 >> (4 * (r - 8))) & 0xFu);
    const unsigned long long bits = (unsigned long long)lo | ((unsigned long long)hi << 32);
    const unsigned h4 = (unsigned)(bits & (bits >> 1) & (bits >> 2) & (bits >> 3) & (bits >> 4));
    if (r >= 4) ok |= (((h0 & h1 & h2 & h3 & h4) >> lane) & 1u) << (r - 4);
    h0 = h1; h1 = h2; h2 = h3; h3 = h4;
  }
  return ok;
}

// cv::borderInterpolate values for the 1-px ring just outside the frame inside a staged window (see the file header).
// Every warp patches the whole window, and every store writes the element's final value (columns only on rows inside
// the frame, from untouched elements; the two ring rows afterwards, from rows inside the frame), so warps that patch
// and read concurrently see the same data and no CTA barrier is needed.
__device__ __noinline__ void warp_patch(float *sf, const StagePlan &pl, int lane, const WarpAccArgs &a) {
  const bool constant = a.border != SSK_BORDER_REPLICATE && a.border != SSK_BORDER_REFLECT && a.border != SSK_BORDER_REFLECT101;
  const int d = a.border == SSK_BORDER_REFLECT101 ? 2 : 1;       // distance from the ring element to its source element
  const float bv = a.bval[0];
  const int cl = -1 - pl.sx0, cr = a.src_cols - pl.sx0;          // window columns of frame columns -1 and src_cols
  const int rt = -1 - pl.sy0, rb = a.src_rows - pl.sy0;          // window rows of frame rows -1 and src_rows
  for (int r = lane; r < GSH; r += 32) {
    float *row = sf + r * WD;
    if ((unsigned)(pl.sy0 + r) >= (unsigned)a.src_rows) continue;
    if (cl >= 0 && cl + d < WD) row[cl] = constant ? bv : row[cl + d];
    if (cr < WD && cr - d >= 0) row[cr] = constant ? bv : row[cr - d];
  }
  __syncwarp();
  for (int c = lane; c < WD; c += 32) {
    if (rt >= 0 && rt + d < GSH) sf[rt * WD + c] = constant ? bv : sf[(rt + d) * WD + c];
    if (rb < GSH && rb - d >= 0) sf[rb * WD + c] = constant ? bv : sf[(rb - d) * WD + c];
  }
  __syncwarp();
}

template <int INTERP, bool WEIGHTS, int MT>
__global__ void __launch_bounds__(TW * NW, 6) k_fused_tma(const __grid_constant__ WarpAccArgs a, const __grid_constant__ Tables tab,
                                                         const TileList tl) {
  typedef StageGeom<SSK_32F> G;
  constexpr int N = Taps<INTERP>::N;
  __shared__ float s_acc[TH][TW];                  // running mean of the tile (on chip for the whole batch)
  __shared__ float s_w[TH][TW];                    // running weight sum of the tile
  __shared__ float4 s_cubic[kInterTab];
  __shared__ __align__(128) unsigned char s_f[2][WIN_BYTES];
  __shared__ __align__(128) float s_g[WEIGHTS ? 2 : 1][WEIGHTS ? GSH * WD : 4];
  __shared__ __align__(8) unsigned long long s_full[2];   // transaction barrier per buffer
  __shared__ int s_cnt[2];                                // warps done with the buffer's current frame
  __shared__ int s_nstaged;
  __shared__ PackedPlan s_plan[KPLAN];
  __shared__ short s_list[KPLAN];                         // frame index of the k-th staged frame
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // border tiles first: their frames cost more (mask flags), so they should not start last
  int tx, ty;
  {
    const int nring = 2 * tl.ntx + 2 * (tl.nty - 2);
    if ((int)blockIdx.x < nring) tile_of_block(tl, blockIdx.x, tx, ty);
    else { const int b = blockIdx.x - nring; tx = 1 + b % (tl.ntx - 2); ty = 1 + b / (tl.ntx - 2); }
  }
  const int bx0 = tx * TW, by0 = ty * TH;
  const int x = bx0 + lane, y0 = by0 + warp * GR;
  const int tw = min(TW, a.cols - bx0), th = min(TH, a.rows - by0);
  const int nrw = max(0, min(GR, th - warp * GR));      // rows of this warp's strip inside the image
  const int nrow = lane < tw ? nrw : 0;
  if (INTERP == SSK_INTER_CUBIC && threadIdx.x < kInterTab) s_cubic[threadIdx.x] = tab.cubic[threadIdx.x];

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
  for (int jj = threadIdx.x; jj < a.njobs; jj += blockDim.x) {
    PackedPlan pp = {0, 0, 0, -1};
    if (a.jobs[jj].ok) {
      StagePlan p = plan_tma(a.jobs[jj].map, bx0, by0, a);
      if (WEIGHTS && !a.jobs[jj].weights) p.staged = 0;    // flat frame (no weight map): generic path
      pp.sx0 = (short)p.sx0; pp.sy0 = (short)p.sy0; pp.sxw = (short)p.sxw; pp.staged = (short)p.staged;
    }
    s_plan[jj] = pp;
    // The tensor maps live in global memory and are rewritten by the host between launches: one acquire fence per CTA
    // and map before its first use (the CTA barrier below orders it before the copies any thread issues later)
    if (pp.staged > 0) {
      tmap_acquire(static_cast<const char *>(a.tmap_frames) + (size_t)jj * 128);
      if (WEIGHTS) tmap_acquire(static_cast<const char *>(a.tmap_weights) + (size_t)jj * 128);
    }
  }
  const unsigned full0 = (unsigned)__cvta_generic_to_shared(&s_full[0]);
  if (threadIdx.x == 0) {
    mbar_init(full0, 1);
    mbar_init(full0 + 8, 1);
    s_cnt[0] = s_cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {       // ordered list of the staged frames
    int base = 0;
    for (int j0 = 0; j0 < a.njobs; j0 += 32) {
      const int jj = j0 + lane;
      const bool st = jj < a.njobs && s_plan[jj].staged > 0;
      const unsigned mk = __ballot_sync(0xffffffffu, st);
      if (st) s_list[base + __popc(mk & ((1u << lane) - 1u))] = (short)jj;
      base += __popc(mk);
    }
    if (lane == 0) s_nstaged = base;
  }
  __syncthreads();
  const int nstaged = s_nstaged;
  // one thread: arm buffer b's barrier and ask the copy engine for the windows of the q-th staged frame
  auto issue = [&](int q, int b) {
    const int jq = s_list[q];
    const PackedPlan pq = s_plan[jq];
    const unsigned mb = full0 + 8 * b;
    const char *tf = static_cast<const char *>(a.tmap_frames) + (size_t)jq * 128;
    mbar_expect_tx(mb, WEIGHTS ? 2 * WIN_BYTES : WIN_BYTES);
    tma_load_2d((unsigned)__cvta_generic_to_shared(s_f[b]), tf, pq.sx0, pq.sy0, mb);
    if (WEIGHTS) {
      const char *tg = static_cast<const char *>(a.tmap_weights) + (size_t)jq * 128;
      tma_load_2d((unsigned)__cvta_generic_to_shared(s_g[b]), tg, pq.sx0, pq.sy0, mb);
    }
    if (q + 1 < nstaged) {       // the next copy uses other tensor maps (one per frame): start fetching them now
      const int jn = s_list[q + 1];
      tmap_prefetch(static_cast<const char *>(a.tmap_frames) + (size_t)jn * 128);
      if (WEIGHTS) tmap_prefetch(static_cast<const char *>(a.tmap_weights) + (size_t)jn * 128);
    }
  };
  if (threadIdx.x == 0) {
    if (nstaged > 0) issue(0, 0);
    if (nstaged > 1) issue(1, 1);
  }

  float *s_acc0 = &s_acc[warp * GR][lane], *s_w0 = &s_w[warp * GR][lane];
  const unsigned acc_a = (unsigned)__cvta_generic_to_shared(s_acc0), w_a = (unsigned)__cvta_generic_to_shared(s_w0);
  const unsigned cub_a = (unsigned)__cvta_generic_to_shared(s_cubic);
  const float xf = (float)min(x, a.cols - 1);            // lanes right of the image follow the last column (never stored)
  const float y0f = (float)y0;
  const bool patch_border = INTERP == SSK_INTER_CUBIC && !(a.border == SSK_BORDER_CONSTANT && a.bval[0] == 0.f);
  // Map coefficients of the staged frames travel lane-wise: lane l holds c[l] of the NEXT staged frame (loaded while the
  // current one is interpolated) and hands it to the warp by shuffles, so no global-memory latency sits at the head of a frame.
  constexpr int NC = MT == MAP_TRANSLATION ? 2 : MT == MAP_AFFINE ? 6 : 7;
  auto map_lane = [&](int jj) -> float { return lane < NC ? __ldg(&a.jobs[jj].map.c[lane]) : 0.f; };
  float mlane = nstaged > 0 ? map_lane(s_list[0]) : 0.f;
  int q = 0;                                             // ordinal of the next staged frame
#pragma unroll 1
  for (int j = 0; j < a.njobs; ++j) {
    const PackedPlan pp = s_plan[j];
    if (pp.staged < 0) continue;
    if (pp.staged == 0) {
      // this (tile, frame) pair cannot be staged (oversize footprint, flat frame, BORDER_WRAP): generic per-pixel path
#pragma unroll 1
      for (int k = 0; k < nrow; ++k) generic_pixel(a, tab, a.jobs[j], x, y0 + k, s_acc0 + k * TW, s_w0 + k * TW);
      continue;
    }
    const int buf = q & 1;
    StagePlan plan; plan.staged = pp.staged; plan.sx0 = pp.sx0; plan.sy0 = pp.sy0; plan.sxw = pp.sxw;
    MapCoef m;
    m.type = MT;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.c[i] = i < NC ? __shfl_sync(0xffffffffu, mlane, i) : 0.f;
    if (q + 1 < nstaged) mlane = map_lane(s_list[q + 1]);
    unsigned okbits = 0xFFu;
    if (nrw > 0 && pp.staged == 1) okbits = warp_okbits<INTERP, MT>(m, bx0, y0, lane, a, tab.cubic_itab);   // overlaps the copy
    okbits &= (1u << nrow) - 1u;
    // every warp waits for the buffer, also one without rows in the image: its arrival below must not run ahead of
    // the frame the buffer holds
    mbar_wait(full0 + 8 * buf, (unsigned)(q >> 1) & 1u);
    if (nrw > 0) {
      float *sfw = reinterpret_cast<float *>(s_f[buf]);
      bool patched = false;
      if (patch_border && pp.staged == 1 &&
          (plan.sx0 < 0 || plan.sy0 < 0 || plan.sx0 + WD > a.src_cols || plan.sy0 + GSH > a.src_rows)) {
        warp_patch(sfw, plan, lane, a);
        patched = true;
      }
      const ColMap<MT> cm(m, xf);
      const unsigned char *sf = s_f[buf];
      const float *sg = s_g[WEIGHTS ? buf : 0];
      constexpr bool C2 = INTERP == SSK_INTER_CUBIC && WEIGHTS;   // packed (frame, weight) bicubic
      if (C2) {
        RollC2 R2;
        R2.ix = INT_MIN; R2.iy = INT_MIN; R2.pf = sf; R2.pw = sg;
        if (nrw == GR) {
          // complete strip: straight-line code, row offsets are immediates
#pragma unroll
          for (int k = 0; k < GR; k += 4) {
            roll_pixel_c2<SSK_32F, MT, 0>(R2, cm, y0f + (float)k, sf, sg, a.scale, plan, cub_a, acc_a + k * TW * 4, w_a + k * TW * 4, (okbits >> k) & 1u);
            roll_pixel_c2<SSK_32F, MT, 1>(R2, cm, y0f + (float)(k + 1), sf, sg, a.scale, plan, cub_a, acc_a + (k + 1) * TW * 4, w_a + (k + 1) * TW * 4, (okbits >> (k + 1)) & 1u);
            roll_pixel_c2<SSK_32F, MT, 2>(R2, cm, y0f + (float)(k + 2), sf, sg, a.scale, plan, cub_a, acc_a + (k + 2) * TW * 4, w_a + (k + 2) * TW * 4, (okbits >> (k + 2)) & 1u);
            roll_pixel_c2<SSK_32F, MT, 3>(R2, cm, y0f + (float)(k + 3), sf, sg, a.scale, plan, cub_a, acc_a + (k + 3) * TW * 4, w_a + (k + 3) * TW * 4, (okbits >> (k + 3)) & 1u);
          }
        } else {
          // strip cut by the bottom edge of the image (last tile row only): one row at a time, window re-anchored per row
#pragma unroll 1
          for (int k = 0; k < nrw; ++k) {
            R2.ix = INT_MIN;
            roll_pixel_c2<SSK_32F, MT, 0>(R2, cm, y0f + (float)k, sf, sg, a.scale, plan, cub_a, acc_a + k * TW * 4, w_a + k * TW * 4, (okbits >> k) & 1u);
          }
        }
      } else {
        RollS<INTERP> R;
        R.ix = INT_MIN; R.iy = INT_MIN; R.pf = sf; R.pw = sg;
#pragma unroll 1
        for (int k = 0; k < GR; k += N) {
          if (k < nrw) roll_pixel_s<SSK_32F, INTERP, WEIGHTS, MT, 0>(R, cm, y0f + (float)k, sf, sg, a.scale, plan, s_cubic, s_acc0 + k * TW, s_w0 + k * TW, (okbits >> k) & 1u);
          if (N > 1 && k + 1 < nrw) roll_pixel_s<SSK_32F, INTERP, WEIGHTS, MT, 1 % N>(R, cm, y0f + (float)(k + 1), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 1) * TW, s_w0 + (k + 1) * TW, (okbits >> (k + 1)) & 1u);
          if (N > 2 && k + 2 < nrw) roll_pixel_s<SSK_32F, INTERP, WEIGHTS, MT, 2 % N>(R, cm, y0f + (float)(k + 2), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 2) * TW, s_w0 + (k + 2) * TW, (okbits >> (k + 2)) & 1u);
          if (N > 3 && k + 3 < nrw) roll_pixel_s<SSK_32F, INTERP, WEIGHTS, MT, 3 % N>(R, cm, y0f + (float)(k + 3), sf, sg, a.scale, plan, s_cubic, s_acc0 + (k + 3) * TW, s_w0 + (k + 3) * TW, (okbits >> (k + 3)) & 1u);
        }
      }
      if (patched) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // patch stores before the buffer's next copy
    }
    // release the buffer: the last warp to arrive re-arms it with the frame two steps ahead
    __syncwarp();
    if (lane == 0) {
      const int old = atomicAdd(&s_cnt[buf], 1);
      if (old == NW - 1) {
        s_cnt[buf] = 0;
        if (q + 2 < nstaged) issue(q + 2, buf);
      }
    }
    ++q;
  }
  __syncthreads();

  if (vec) {
    for (int k = threadIdx.x; k < th * (TW / 4); k += blockDim.x) {
      const int r = k / (TW / 4), q4 = k - r * (TW / 4);
      *reinterpret_cast<float4 *>(a.acc + (int64_t)(by0 + r) * a.cols + bx0 + 4 * q4) = reinterpret_cast<const float4 *>(s_acc[r])[q4];
      *reinterpret_cast<float4 *>(a.wacc + (int64_t)(by0 + r) * a.cols + bx0 + 4 * q4) = reinterpret_cast<const float4 *>(s_w[r])[q4];
    }
  } else {
    for (int k = threadIdx.x; k < th * tw; k += blockDim.x) {
      const int r = k / tw, q4 = k - r * tw;
      a.acc[(int64_t)(by0 + r) * a.cols + bx0 + q4] = s_acc[r][q4];
      a.wacc[(int64_t)(by0 + r) * a.cols + bx0 + q4] = s_w[r][q4];
    }
  }
}

template <int INTERP, bool WEIGHTS>
void launch_tma_mt(const WarpAccArgs &a, const Tables &tab, const TileList &tl, cudaStream_t s) {
  const int nt = tl.ntx * tl.nty;
  if (a.map_type == MAP_AFFINE) k_fused_tma<INTERP, WEIGHTS, MAP_AFFINE><<<nt, TW * NW, 0, s>>>(a, tab, tl);
  else if (a.map_type == MAP_TRANSLATION) k_fused_tma<INTERP, WEIGHTS, MAP_TRANSLATION><<<nt, TW * NW, 0, s>>>(a, tab, tl);
  else k_fused_tma<INTERP, WEIGHTS, MAP_EUCLIDEAN><<<nt, TW * NW, 0, s>>>(a, tab, tl);
}

}  // namespace

bool fused_tma_applicable(const WarpAccArgs &a) {
  if (a.depth != SSK_32F || a.cn != 1 || !a.tmap_frames || (a.use_weights && !a.tmap_weights)) return false;
  if (!(a.map_type == MAP_AFFINE || a.map_type == MAP_TRANSLATION || a.map_type == MAP_EUCLIDEAN)) return false;
  if (a.border == SSK_BORDER_WRAP) return false;
  const int ntx = div_up(a.cols, TW), nty = div_up(a.rows, TH);
  return ntx >= 3 && nty >= 3 && a.src_cols < 30000 && a.src_rows < 30000;
}

// One launch per KPLAN frames over every tile of the accumulator.  Requires fused_tma_applicable(a).
int launch_warp_accumulate_tma(const WarpAccArgs &a_in, const Tables &tab, cudaStream_t s) {
  WarpAccArgs a = a_in;
  TileList tl;
  tl.ntx = div_up(a.cols, TW); tl.nty = div_up(a.rows, TH); tl.ring = 1;
  const FrameJob *jobs = a.jobs;
  const int njobs = a.njobs;
  for (int j0 = 0; j0 < njobs; j0 += KPLAN) {
    a.jobs = jobs + j0; a.njobs = std::min(KPLAN, njobs - j0);
    a.tmap_frames = static_cast<const char *>(a_in.tmap_frames) + (size_t)j0 * 128;
    a.tmap_weights = a_in.tmap_weights ? static_cast<const char *>(a_in.tmap_weights) + (size_t)j0 * 128 : nullptr;
    if (a.use_weights) {
      if (a.interp == SSK_INTER_CUBIC) launch_tma_mt<SSK_INTER_CUBIC, true>(a, tab, tl, s);
      else if (a.interp == SSK_INTER_NEAREST) launch_tma_mt<SSK_INTER_NEAREST, true>(a, tab, tl, s);
      else launch_tma_mt<SSK_INTER_LINEAR, true>(a, tab, tl, s);
    } else {
      if (a.interp == SSK_INTER_CUBIC) launch_tma_mt<SSK_INTER_CUBIC, false>(a, tab, tl, s);
      else if (a.interp == SSK_INTER_NEAREST) launch_tma_mt<SSK_INTER_NEAREST, false>(a, tab, tl, s);
      else launch_tma_mt<SSK_INTER_LINEAR, false>(a, tab, tl, s);
    }
    SSK_LAUNCH_CHECK();
  }
  return SSK_OK;
}

}  // namespace ssk
