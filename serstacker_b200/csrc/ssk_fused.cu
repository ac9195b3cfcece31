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
#include "ssk_fused_impl.cuh"

namespace ssk {

namespace {


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
    generic_pixel_fast(a, tab, a.jobs[j], x, y, A, &W);
  }
  a.wacc[p] = W;
  for (int c = 0; c < a.cn; ++c) a.acc[p * a.cn + c] = A[c];
}

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
  SSK_REQUIRE(a.interp == SSK_INTER_NEAREST || a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC || a.interp == SSK_INTER_LANCZOS4,
              "warp_accumulate: interpolation must be NEAREST, LINEAR, CUBIC or LANCZOS4");
  SSK_REQUIRE(a.border != SSK_BORDER_TRANSPARENT, "warp_accumulate: BORDER_TRANSPARENT is not meaningful here");
  SSK_REQUIRE(a.cn >= 1 && a.cn <= 4, "warp_accumulate: 1..4 channels");
  SSK_REQUIRE(a.depth == SSK_32F || a.depth == SSK_16U || a.depth == SSK_8U, "warp_accumulate: unsupported frame depth");
  a.stage_aligned = a.stage_aligned && (a.src_step % 16 == 0) && (a.w_step % 16 == 0);
  const bool lanczos = a.interp == SSK_INTER_LANCZOS4;     // 8 x 8 taps: the one-thread-per-pixel form only
  if (!lanczos && fused_tma_applicable(a)) return launch_warp_accumulate_tma(a, tab, s);
  const int ntx = div_up(a.cols, TW), nty = div_up(a.rows, TH);
  const bool staged = !lanczos && a.cn == 1 && ntx >= 3 && nty >= 3 && a.src_cols < 32000 && a.src_rows < 32000 &&
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
