// K5b: the Bayer form of the fused per-frame loop.  For a batch of raw Bayer frames, one launch gathers, per pixel of
// the accumulator and per frame, the four raw samples around the frame's remap position into the per-colour sums
// and counters, under the eroded validity mask of the frame's remap.  acc / cntr are read and written once per batch.
//
// Reference semantics reproduced here:
//   c_image_stacking_pipeline::process_input_sequence, bayer branch
//        core/pipeline/c_image_stacking_pipeline/c_image_stacking_pipeline.cc:1644-1651 (custom_remap -> current_mask),
//        :1730-1752 (set_bayer_pattern, set_remap(current_remap), add(_raw_bayer_image, current_mask))
//   c_frame_registration::base_remap mask: erode5x5(remap(all-255, interp, CONSTANT 0) >= 255), border value 255
//        core/proc/image_registration/c_frame_registration.cc:1315-1337
//   _bayer_accumulate, rmap + CV_8UC1 mask branch   core/average/c_frame_accumulation.cc:1040-1075, 1097-1110
//        src_x = (int)p[0]; ax = (src_x + 1 - p[0]) evaluated in float then widened; s = ax * ay * w in double;
//        acc[c] += src * s and cntr[c] += s: float += double (sum formed in double, narrowed once), taps in the
//        order 00, 01, 10, 11 (the two green taps of a 2x2 cell update the same channel one after the other).
#include <cmath>
#include "ssk_warp.cuh"

namespace ssk {
namespace {

constexpr int BTW = 32, BTH = 8;      // pixels per CTA (one pixel per thread)
constexpr int BPLAN = 256;            // frames per launch (per-CTA table of tile flags)

// current_remap at accumulator pixel (x, y): the analytic map, or flow + grid when c_eccflow refined it
// (ecc_flow_to_remap, c_frame_registration.cc:900-917; same form as k_fused_flow's map_at)
__device__ __forceinline__ void remap_at(const MapCoef &m, const float2 *flow, int x, int y, int src_cols, float &u, float &v) {
  if (flow) {
    const float2 d = __ldg(flow + (int64_t)y * src_cols + x);
    u = __fadd_rn(d.x, (float)x); v = __fadd_rn(d.y, (float)y);
  } else {
    map_xy(m, (float)x, (float)y, u, v);
  }
}

// pre-erosion flag of base_remap's mask at accumulator pixel (x, y)
__device__ __forceinline__ bool flag_at(const MapCoef &m, const float2 *flow, int interp, int x, int y, int src_cols, int src_rows,
                                        const short *itab) {
  float u, v;
  remap_at(m, flow, x, y, src_cols, u, v);
  return valid255(interp, u, v, src_cols, src_rows, itab);
}

// mask(x, y) of base_remap: 5x5 erosion (border value 255: positions outside the image do not erode)
__device__ __noinline__ bool mask_at(const MapCoef &m, const float2 *flow, int interp, int x, int y, int cols, int rows, int src_cols,
                                     int src_rows, const short *itab) {
#pragma unroll 1
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = y + dy;
    if ((unsigned)yy >= (unsigned)rows) continue;
#pragma unroll 1
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = x + dx;
      if ((unsigned)xx >= (unsigned)cols) continue;
      if (!flag_at(m, flow, interp, xx, yy, src_cols, src_rows, itab)) return false;
    }
  }
  return true;
}

// 1: every pixel of the tile and of its 2-px erosion halo maps into the tap-safe interior of the frame (mask = 255 and
// the four raw samples are in bounds); 0: decide per pixel.  Affine-like maps are monotone along both axes of the
// tile, so the four corners bound the footprint; projective maps always take the per-pixel path.
__device__ __noinline__ int tile_safe(const MapCoef &m, int bx0, int by0, const WarpAccArgs &a) {
  if (m.type == MAP_HOMOGRAPHY) return 0;
  const int cx0 = max(bx0 - 2, 0), cy0 = max(by0 - 2, 0);
  const int cx1 = min(bx0 + BTW + 1, a.cols - 1), cy1 = min(by0 + BTH + 1, a.rows - 1);
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    float u, v;
    map_xy(m, (float)((k & 1) ? cx1 : cx0), (float)((k & 2) ? cy1 : cy0), u, v);
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
  }
  return (umin >= 3.f && vmin >= 3.f && umax <= (float)(a.src_cols - 4) && vmax <= (float)(a.src_rows - 4)) ? 1 : 0;
}

// colour channel (B = 0, G = 1, R = 2; c_frame_accumulation.h:230-234) of the 2x2 cell position q = (y & 1) * 2 + (x & 1),
// packed two bits per position (c_frame_accumulation.cc:1262-1334)
__host__ __device__ inline unsigned pattern_code(int colorid) {
  switch (colorid) {
    case SSK_COLORID_BAYER_RGGB: return 2u | (1u << 2) | (1u << 4) | (0u << 6);
    case SSK_COLORID_BAYER_GRBG: return 1u | (2u << 2) | (0u << 4) | (1u << 6);
    case SSK_COLORID_BAYER_GBRG: return 1u | (0u << 2) | (2u << 4) | (1u << 6);
    default: /* BGGR */          return 0u | (1u << 2) | (1u << 4) | (2u << 6);
  }
}

// (double)i for 0 <= i < 2^31 without a conversion instruction: 2^52 + i is exact in the low mantissa word
__device__ __forceinline__ double int_as_double(int i) { return __hiloint2double(0x43300000, i) - 4503599627370496.0; }

// The reference's tap weights along one axis: a = (double)((float)(s + 1) - u), b = (double)(u - (float)s), float differences
// widened to double.  For s >= 1 (u in [s, s + 1)) both float differences are exact (Sterbenz), so they equal the double
// differences of the widened operands: one widening of u instead of two, and no int -> float conversions.  s = 0 (u may be
// negative or tiny: 1 - u rounds in float) keeps the literal form.
__device__ __forceinline__ void tap_weights(float u, int s, double &a, double &b) {
  if (s >= 1) {
    const double du = (double)u, ds = int_as_double(s);
    b = du - ds;
    a = (ds + 1.0) - du;
  } else {
    a = (double)((float)(s + 1) - u);
    b = (double)(u - (float)s);
  }
}

// (double)(sample * scale) as the reference forms it (convertTo(CV_32F, 1 / (1 << bpp)), then widened).  EXACT: the scale is
// a power of two, so the float product of an 8 / 16-bit sample is exact and equals the double product of the exactly
// converted operands - no conversion instruction on the way.
template <int DEPTH, bool EXACT>
__device__ __forceinline__ double sample_d(const Img &im, int y, int x, double scale_d) {
  if (DEPTH == SSK_32F || !EXACT) return (double)load_px<DEPTH>(im, y, x, 0);
  typedef typename PixT<DEPTH>::type T;
  const T *row = reinterpret_cast<const T *>(static_cast<const char *>(im.data) + (int64_t)y * im.step);
  return int_as_double((int)__ldg(row + x)) * scale_d;
}

template <int DEPTH, bool EXACT>
__global__ void __launch_bounds__(BTW * BTH) k_fused_bayer(const __grid_constant__ WarpAccArgs a, const __grid_constant__ Tables tab,
                                                          const unsigned pcode, const double scale_d) {
  __shared__ signed char s_flag[BPLAN];
  const int bx0 = blockIdx.x * BTW, by0 = blockIdx.y * BTH;
  for (int jj = threadIdx.x; jj < a.njobs; jj += BTW * BTH)
    s_flag[jj] = a.jobs[jj].ok ? (signed char)(a.flow ? 0 : tile_safe(a.jobs[jj].map, bx0, by0, a)) : (signed char)-1;   // per-pixel maps: decide per pixel
  __syncthreads();
  const int x = bx0 + (threadIdx.x & 31), y = by0 + (threadIdx.x >> 5);
  if (x >= a.cols || y >= a.rows) return;
  const int64_t p = ((int64_t)y * a.cols + x) * 3;
  float A[3] = {a.acc[p], a.acc[p + 1], a.acc[p + 2]};
  float N[3] = {a.wacc[p], a.wacc[p + 1], a.wacc[p + 2]};
  // positions (row parity, column parity) of R and B in the 2x2 cell, column parity of G on even rows (frame sizes are even,
  // so the pattern table of c_frame_accumulation.cc:1262-1334 covers every sample)
  int rR = 0, cR = 0, rB = 0, cB = 0, cg0 = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int cc = (int)((pcode >> (2 * q)) & 3u);
    if (cc == 2) { rR = q >> 1; cR = q & 1; }
    if (cc == 0) { rB = q >> 1; cB = q & 1; }
    if (cc == 1 && (q >> 1) == 0) cg0 = q & 1;
  }
  Img im;
  im.data = nullptr; im.step = a.src_step; im.rows = a.src_rows; im.cols = a.src_cols; im.depth = DEPTH; im.cn = 1; im.scale = a.scale;
#pragma unroll 1
  for (int j = 0; j < a.njobs; ++j) {
    const int flag = s_flag[j];
    if (flag < 0) continue;                       // frame dropped by the registration
    const MapCoef &m = a.jobs[j].map;
    const float2 *flow = a.flow ? a.flow + (int64_t)j * a.flow_stride : nullptr;
    if (!flag && !mask_at(m, flow, a.interp, x, y, a.cols, a.rows, a.src_cols, a.src_rows, tab.cubic_itab)) continue;
    float u, v;
    remap_at(m, flow, x, y, a.src_cols, u, v);
    const int sx = (int)u, sy = (int)v;           // truncation toward zero, as the reference's (int) cast
    if (!(sx >= 0 && sx < a.src_cols - 1 && sy >= 0 && sy < a.src_rows - 1)) continue;
    double ax, bx, ay, by;
    tap_weights(u, sx, ax, bx);
    tap_weights(v, sy, ay, by);
    im.data = a.jobs[j].frame;
    // The 2x2 footprint holds one R, one B and two G samples whatever the parity of (sx, sy).  The reference walks the taps
    // in the order 00, 01, 10, 11 and updates the channel of each; channels do not interact, so only the order of the two G
    // updates matters: the G of the top row first.  Walking by channel instead of by tap keeps the updates free of per-lane
    // channel selects (the tap-order form compiled to three predicated copies of every conversion).
    auto update = [&](float &Ac, float &Nc, int dy, int dx) {
      const double w = (dx ? bx : ax) * (dy ? by : ay);
      const double sv = sample_d<DEPTH, EXACT>(im, sy + dy, sx + dx, scale_d);
      const double an = (double)Ac + sv * w;
      const double nn = (double)Nc + w;
      Ac = (float)an; Nc = (float)nn;
    };
    const int gdx = (sx ^ cg0 ^ sy) & 1;                      // column offset of the G sample in row sy
    update(A[1], N[1], 0, gdx);
    update(A[1], N[1], 1, gdx ^ 1);
    update(A[2], N[2], (sy ^ rR) & 1, (sx ^ cR) & 1);
    update(A[0], N[0], (sy ^ rB) & 1, (sx ^ cB) & 1);
  }
  a.acc[p] = A[0]; a.acc[p + 1] = A[1]; a.acc[p + 2] = A[2];
  a.wacc[p] = N[0]; a.wacc[p + 1] = N[1]; a.wacc[p + 2] = N[2];
}

}  // namespace

// a.acc / a.wacc: rows x cols x 3 sums and counters (c_bayer_average::_accumulator / _counter); a.jobs[j].frame: raw Bayer
// frames (one channel).  a.interp selects the interpolation of the validity mask (registration_options.interpolation).
int launch_bayer_warp_accumulate(const WarpAccArgs &a_in, const Tables &tab, int colorid, cudaStream_t s) {
  WarpAccArgs a = a_in;
  SSK_REQUIRE(a.interp == SSK_INTER_NEAREST || a.interp == SSK_INTER_LINEAR || a.interp == SSK_INTER_CUBIC,
              "bayer_warp_accumulate: interpolation must be NEAREST, LINEAR or CUBIC");
  SSK_REQUIRE(a.cn == 1, "bayer_warp_accumulate: raw Bayer frames have one channel");
  SSK_REQUIRE(a.rows == a.src_rows && a.cols == a.src_cols, "bayer_warp_accumulate: accumulator and frame sizes differ");
  SSK_REQUIRE(!(a.src_rows & 1) && !(a.src_cols & 1), "bayer_warp_accumulate: frame size must be even");
  SSK_REQUIRE(colorid >= SSK_COLORID_BAYER_RGGB && colorid <= SSK_COLORID_BAYER_BGGR, "bayer_warp_accumulate: RGGB/GRBG/GBRG/BGGR");
  const unsigned pcode = pattern_code(colorid);
  int sexp = 0;
  const bool exact = a.scale > 0.f && std::frexp(a.scale, &sexp) == 0.5f && sexp > -100 && !getenv("SSK_BAYER_LITERAL");   // 1 / (1 << bpp), or 1
  const dim3 grid(div_up(a.cols, BTW), div_up(a.rows, BTH));
  const FrameJob *jobs = a.jobs;
  const int njobs = a.njobs;
  for (int j0 = 0; j0 < njobs; j0 += BPLAN) {
    a.jobs = jobs + j0; a.njobs = std::min(BPLAN, njobs - j0);
    if (a_in.flow) a.flow = a_in.flow + (int64_t)j0 * a.flow_stride;
    if (a.depth == SSK_32F) k_fused_bayer<SSK_32F, false><<<grid, BTW * BTH, 0, s>>>(a, tab, pcode, (double)a.scale);
    else if (a.depth == SSK_16U && exact) k_fused_bayer<SSK_16U, true><<<grid, BTW * BTH, 0, s>>>(a, tab, pcode, (double)a.scale);
    else if (a.depth == SSK_16U) k_fused_bayer<SSK_16U, false><<<grid, BTW * BTH, 0, s>>>(a, tab, pcode, (double)a.scale);
    else if (a.depth == SSK_8U && exact) k_fused_bayer<SSK_8U, true><<<grid, BTW * BTH, 0, s>>>(a, tab, pcode, (double)a.scale);
    else if (a.depth == SSK_8U) k_fused_bayer<SSK_8U, false><<<grid, BTW * BTH, 0, s>>>(a, tab, pcode, (double)a.scale);
    else { set_error("bayer_warp_accumulate: unsupported frame depth"); return SSK_ERR_INVALID; }
    SSK_LAUNCH_CHECK();
  }
  return SSK_OK;
}

}  // namespace ssk
