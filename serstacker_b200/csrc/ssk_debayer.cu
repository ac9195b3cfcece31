// debayer_nn2 (core/io/debayer.cc:827-1195): the bilinear demosaic read_input_frame applies to raw Bayer frames
// (c_image_stacking_pipeline_base.cc:125-279).  One thread per raw pixel gathers its 3x3 neighbourhood (rows / columns
// -1 -> 1 and N -> N-2, as the reference indexes its first / last row and pixel pair) and writes the BGR triple of the
// source depth: integer depths round as (n/2 + sum) / n, CV_32F sums left to right in the reference's operand order
// (diagonals: TL + TR + BL + BR; cross: T + L + R + B).  HBM-bound: 1 sample read (+ L1/L2-served neighbours), 3 written.
#include "ssk_prep.cuh"

namespace ssk {
namespace {

template <class T> struct Wide { typedef unsigned type; };   // samples and their sums are non-negative: / 4, / 2 are shifts
template <> struct Wide<float> { typedef float type; };

template <class T> __device__ __forceinline__ T avg4(typename Wide<T>::type a, typename Wide<T>::type b, typename Wide<T>::type c,
                                                      typename Wide<T>::type d) { return (T)((2 + a + b + c + d) / 4); }
template <> __device__ __forceinline__ float avg4<float>(float a, float b, float c, float d) {
  return __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 4.0f);
}
template <class T> __device__ __forceinline__ T avg2(typename Wide<T>::type a, typename Wide<T>::type b) { return (T)((1 + a + b) / 2); }
template <> __device__ __forceinline__ float avg2<float>(float a, float b) { return __fdiv_rn(__fadd_rn(a, b), 2.0f); }

template <class T>
__global__ void __launch_bounds__(256) k_debayer_nn2(const T *__restrict__ src, int64_t sstep, int rows, int cols, int ry, int rx,
                                                     T *__restrict__ dst, int64_t dstep) {
  typedef typename Wide<T>::type W;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int ym = y == 0 ? 1 : y - 1, yp = y == rows - 1 ? rows - 2 : y + 1;
  const int xm = x == 0 ? 1 : x - 1, xp = x == cols - 1 ? cols - 2 : x + 1;
  const T *s0 = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)ym * sstep);
  const T *s1 = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)y * sstep);
  const T *s2 = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)yp * sstep);
  const W c = (W)__ldg(s1 + x);
  const int py = y & 1, px = x & 1;
  const bool is_r = py == ry && px == rx, is_b = py != ry && px != rx;
  T r, g, b;
  if (is_r || is_b) {
    const T diag = avg4<T>((W)__ldg(s0 + xm), (W)__ldg(s0 + xp), (W)__ldg(s2 + xm), (W)__ldg(s2 + xp));
    g = avg4<T>((W)__ldg(s0 + x), (W)__ldg(s1 + xm), (W)__ldg(s1 + xp), (W)__ldg(s2 + x));
    r = is_r ? (T)c : diag;
    b = is_r ? diag : (T)c;
  } else {
    const T vert = avg2<T>((W)__ldg(s0 + x), (W)__ldg(s2 + x));
    const T horz = avg2<T>((W)__ldg(s1 + xm), (W)__ldg(s1 + xp));
    g = (T)c;
    const bool on_r_row = py == ry;
    r = on_r_row ? horz : vert;
    b = on_r_row ? vert : horz;
  }
  T *o = reinterpret_cast<T *>(reinterpret_cast<char *>(dst) + (int64_t)y * dstep) + (int64_t)x * 3;
  o[0] = b; o[1] = g; o[2] = r;
}

// N raw samples per thread (N = 4, or 16 bytes' worth): one vector load per source row (+ the two neighbours), three vector
// stores of the 3N output samples.  Same arithmetic as k_debayer_nn2; needs cols % N == 0 and rows / pointers aligned to N
// samples.
template <class T, int N> struct alignas(N * sizeof(T)) Pack { T v[N]; };

template <class T, int N>
__global__ void __launch_bounds__(256) k_debayer_nn2_vec(const T *__restrict__ src, int64_t sstep, int rows, int cols, int ry, int rx,
                                                        T *__restrict__ dst, int64_t dstep) {
  typedef typename Wide<T>::type W;
  const int x0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * N, y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x0 >= cols || y >= rows) return;
  const int yr[3] = {y == 0 ? 1 : y - 1, y, y == rows - 1 ? rows - 2 : y + 1};
  W w[3][N + 2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const T *row = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)yr[r] * sstep);
    const Pack<T, N> p = *reinterpret_cast<const Pack<T, N> *>(row + x0);
#pragma unroll
    for (int i = 0; i < N; ++i) w[r][i + 1] = (W)p.v[i];
    w[r][0] = x0 == 0 ? (W)p.v[1] : (W)__ldg(row + x0 - 1);                   // column -1 -> 1
    w[r][N + 1] = x0 + N == cols ? (W)p.v[N - 2] : (W)__ldg(row + x0 + N);    // column W -> W - 2
  }
  Pack<T, N> o[3];
  T *ov = &o[0].v[0];
  const int py = y & 1;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int px = i & 1;                                                   // x0 is even
    const bool is_r = py == ry && px == rx, is_b = py != ry && px != rx;
    const W c = w[1][i + 1];
    T r, g, b;
    if (is_r || is_b) {
      const T diag = avg4<T>(w[0][i], w[0][i + 2], w[2][i], w[2][i + 2]);
      g = avg4<T>(w[0][i + 1], w[1][i], w[1][i + 2], w[2][i + 1]);
      r = is_r ? (T)c : diag;
      b = is_r ? diag : (T)c;
    } else {
      const T vert = avg2<T>(w[0][i + 1], w[2][i + 1]);
      const T horz = avg2<T>(w[1][i], w[1][i + 2]);
      g = (T)c;
      const bool on_r_row = py == ry;
      r = on_r_row ? horz : vert;
      b = on_r_row ? vert : horz;
    }
    ov[3 * i] = b; ov[3 * i + 1] = g; ov[3 * i + 2] = r;
  }
  Pack<T, N> *out = reinterpret_cast<Pack<T, N> *>(reinterpret_cast<T *>(reinterpret_cast<char *>(dst) + (int64_t)y * dstep) + (int64_t)x0 * 3);
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
}

template <class T, int N>
bool try_debayer_vec(const void *src, int64_t sstep, int rows, int cols, int ry, int rx, void *dst, int64_t dstep, cudaStream_t s) {
  const size_t a = N * sizeof(T);
  if (!(cols % N == 0 && reinterpret_cast<uintptr_t>(src) % a == 0 && reinterpret_cast<uintptr_t>(dst) % a == 0 &&
        sstep % (int64_t)a == 0 && dstep % (int64_t)a == 0)) return false;
  dim3 grid(div_up(cols / N, 32), div_up(rows, 8));
  k_debayer_nn2_vec<T, N><<<grid, 256, 0, s>>>(static_cast<const T *>(src), sstep, rows, cols, ry, rx, static_cast<T *>(dst), dstep);
  return true;
}

template <class T>
void run_debayer(const void *src, int64_t sstep, int rows, int cols, int ry, int rx, void *dst, int64_t dstep, cudaStream_t s) {
  constexpr int NW = 16 / (int)sizeof(T);        // 16-byte vectors
  if (NW > 4 && try_debayer_vec<T, NW>(src, sstep, rows, cols, ry, rx, dst, dstep, s)) return;
  if (try_debayer_vec<T, 4>(src, sstep, rows, cols, ry, rx, dst, dstep, s)) return;
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_debayer_nn2<T><<<grid, 256, 0, s>>>(static_cast<const T *>(src), sstep, rows, cols, ry, rx, static_cast<T *>(dst), dstep);
}


// ---- debayer_nn2 -> BGR2GRAY -> cv::pyrDown in one pass ------------------------------------------------------------
// The registration side of the bayer_average loop (read_input_frame's debayer, extract_channel gray, scaleImage's pyrDown at
// ecc.scale 0.5: c_image_stacking_pipeline_base.cc:221-236, c_frame_registration.cc:203-250) only needs the half-size gray
// ECC image, so the demosaiced BGR frame (3 samples written and read back per raw sample) is never formed: a CTA demosaics the
// gray window of its 64 x 16 output tile into shared memory - one thread per 2 x 2 Bayer cell, the cell's 4 x 4 raw neighbourhood
// in registers - and applies pyrDown's 5 x 5 from there.  Same arithmetic as k_debayer_nn2 -> load_gray -> k_pyrdown (integer
// rounding of the demosaic, sample * 1/(1 << bpp), the gray FMAs, pyrDown's operand orders), so the result is bit-identical.
constexpr int BP_OW = 64, BP_OH = 16, BP_RPT = 4;
constexpr int BP_UH = BP_OH + 2, BP_UW = BP_OW / 2 + 2;        // units of 2 rows x 4 columns covering the gray window
constexpr int BP_GH = 2 * BP_UH, BP_GW = 4 * BP_UW;            // rows 2 oy0 - 2 ..., columns 2 ox0 - 4 ... (4-sample aligned)
constexpr int BP_NIT = (BP_UH * BP_UW + 255) / 256;

template <class T> __device__ __forceinline__ float sample_f(T v, float scale) { return __fmul_rn((float)v, scale); }
template <> __device__ __forceinline__ float sample_f<float>(float v, float) { return v; }

// BGR of the pixel at the centre of the 3 x 3 window w (rows / columns already border-mapped), parities (py, px)
template <class T>
__device__ __forceinline__ float debayer_gray(const typename Wide<T>::type (&w)[3][3], int py, int px, int ry, int rx, float scale) {
  const bool is_r = py == ry && px == rx, is_b = py != ry && px != rx;
  T r, g, b;
  if (is_r || is_b) {
    const T diag = avg4<T>(w[0][0], w[0][2], w[2][0], w[2][2]);
    g = avg4<T>(w[0][1], w[1][0], w[1][2], w[2][1]);
    r = is_r ? (T)w[1][1] : diag;
    b = is_r ? diag : (T)w[1][1];
  } else {
    const T vert = avg2<T>(w[0][1], w[2][1]);
    const T horz = avg2<T>(w[1][0], w[1][2]);
    g = (T)w[1][1];
    const bool on_r_row = py == ry;
    r = on_r_row ? horz : vert;
    b = on_r_row ? vert : horz;
  }
  return bgr2gray(sample_f<T>(b, scale), sample_f<T>(g, scale), sample_f<T>(r, scale));
}

template <class T, bool VEC, int RY, int RX>      // (RY, RX): parities of the R sample of the 2 x 2 cell, compile-time
__global__ void __launch_bounds__(256) k_bayer_gray_pyrdown(const void *const *__restrict__ src_ptrs, int64_t sstep, int rows, int cols,
                                                            float scale, float *const *__restrict__ dst_ptrs, int dst_rows, int dst_cols) {
  typedef typename Wide<T>::type W;
  constexpr int ry = RY, rx = RX;
  __shared__ __align__(16) float sg[BP_GH][BP_GW + 4];     // pitch a multiple of 4: 16-byte row stores, 8-byte row loads
  const char *src = static_cast<const char *>(src_ptrs[blockIdx.z]);
  float *dst = dst_ptrs[blockIdx.z];
  const int gx0 = 2 * (int)blockIdx.x * BP_OW - 4, gy0 = 2 * (int)blockIdx.y * BP_OH - 2;   // multiples of 4 / 2
  auto row_ptr = [&](int y) { return reinterpret_cast<const T *>(src + (int64_t)y * sstep); };
  // ---- gray window: one unit of 2 rows x 4 columns (two Bayer cells) per thread and pass, its 4 x 6 raw neighbourhood in
  // registers (VEC: one aligned 4-sample load + the two neighbours per row); the passes are unrolled so that the loads of
  // all of a thread's units are in flight together
#pragma unroll
  for (int it = 0; it < BP_NIT; ++it) {
    const int c = threadIdx.x + it * 256;
    if (c >= BP_UH * BP_UW) break;
    const int uy = c / BP_UW, ux = c - uy * BP_UW;
    const int y0 = gy0 + 2 * uy, x0 = gx0 + 4 * ux;
    if (y0 >= 0 && y0 + 1 < rows && x0 >= 0 && x0 + 3 < cols) {
      // debayer_nn2's own border rule for the neighbourhood: row / column -1 -> 1, N -> N - 2
      const int yy[4] = {y0 == 0 ? 1 : y0 - 1, y0, y0 + 1, y0 + 2 == rows ? rows - 2 : y0 + 2};
      const int xm = x0 == 0 ? 1 : x0 - 1, xp = x0 + 4 == cols ? cols - 2 : x0 + 4;
      W q[4][6];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const T *rp = row_ptr(yy[i]);
        if (VEC) {
          const Pack<T, 4> v = *reinterpret_cast<const Pack<T, 4> *>(rp + x0);
#pragma unroll
          for (int k = 0; k < 4; ++k) q[i][k + 1] = (W)v.v[k];
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) q[i][k + 1] = (W)__ldg(rp + x0 + k);
        }
        q[i][0] = (W)__ldg(rp + xm);
        q[i][5] = (W)__ldg(rp + xp);
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        float gq[4];
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          W w[3][3];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) w[i][k] = q[dy + i][dx + k];
          gq[dx] = debayer_gray<T>(w, dy, dx & 1, ry, rx, scale);      // y0 is even, x0 a multiple of 4
        }
        *reinterpret_cast<float4 *>(&sg[2 * uy + dy][4 * ux]) = make_float4(gq[0], gq[1], gq[2], gq[3]);
      }
    } else {
      // positions outside the frame: cv::pyrDown's BORDER_REFLECT101 of the gray image, pixel by pixel
#pragma unroll 1
      for (int k8 = 0; k8 < 8; ++k8) {
        const int y = border_idx(y0 + (k8 >> 2), rows, SSK_BORDER_REFLECT101), x = border_idx(x0 + (k8 & 3), cols, SSK_BORDER_REFLECT101);
        const int yy[3] = {y == 0 ? 1 : y - 1, y, y == rows - 1 ? rows - 2 : y + 1};
        const int xx[3] = {x == 0 ? 1 : x - 1, x, x == cols - 1 ? cols - 2 : x + 1};
        W w[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int k = 0; k < 3; ++k) w[i][k] = (W)__ldg(row_ptr(yy[i]) + xx[k]);
        sg[2 * uy + (k8 >> 2)][4 * ux + (k8 & 3)] = debayer_gray<T>(w, y & 1, x & 1, ry, rx, scale);
      }
    }
  }
  __syncthreads();
  // ---- pyrDown from the window (k_pyrdown's forms: ssk_prep.cu); gray column 2 ox - 2 sits at window column 2 lx + 2
  const int lx = threadIdx.x & (BP_OW - 1), ly0 = (threadIdx.x / BP_OW) * BP_RPT;
  const int ox = blockIdx.x * BP_OW + lx, oy0 = blockIdx.y * BP_OH + ly0;
  if (ox >= dst_cols || oy0 >= dst_rows) return;
  const int width0 = min((cols - 3) / 2 + 1, dst_cols);
  const bool hsimd = ox >= 1 && ox < 1 + 4 * ((width0 - 1) / 4);
  const bool vsimd = ox < (dst_cols & ~3);
  constexpr int NR = 2 * BP_RPT + 3;
  float h[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float *g = &sg[2 * ly0 + r][2 * lx + 2];
    const float2 g01 = *reinterpret_cast<const float2 *>(g), g23 = *reinterpret_cast<const float2 *>(g + 2);
    h[r] = pd_hform(g01.x, g01.y, g23.x, g23.y, g[4], hsimd);
  }
#pragma unroll
  for (int j = 0; j < BP_RPT; ++j) {
    if (oy0 + j >= dst_rows) break;
    dst[(int64_t)(oy0 + j) * dst_cols + ox] = pd_vform(h[2 * j], h[2 * j + 1], h[2 * j + 2], h[2 * j + 3], h[2 * j + 4], vsimd);
  }
}

template <class T, bool VEC>
void run_bayer_gray_pyrdown_p(const void *const *src_ptrs, int64_t sstep, int rows, int cols, int ry, int rx, float scale,
                              float *const *dst_ptrs, int dr, int dc, int batch, cudaStream_t s) {
  const dim3 grid(div_up(dc, BP_OW), div_up(dr, BP_OH), batch);
  if (ry == 0 && rx == 0) k_bayer_gray_pyrdown<T, VEC, 0, 0><<<grid, 256, 0, s>>>(src_ptrs, sstep, rows, cols, scale, dst_ptrs, dr, dc);
  else if (ry == 0) k_bayer_gray_pyrdown<T, VEC, 0, 1><<<grid, 256, 0, s>>>(src_ptrs, sstep, rows, cols, scale, dst_ptrs, dr, dc);
  else if (rx == 0) k_bayer_gray_pyrdown<T, VEC, 1, 0><<<grid, 256, 0, s>>>(src_ptrs, sstep, rows, cols, scale, dst_ptrs, dr, dc);
  else k_bayer_gray_pyrdown<T, VEC, 1, 1><<<grid, 256, 0, s>>>(src_ptrs, sstep, rows, cols, scale, dst_ptrs, dr, dc);
}

template <class T>
void run_bayer_gray_pyrdown(const void *const *src_ptrs, bool aligned, int64_t sstep, int rows, int cols, int ry, int rx, float scale,
                            float *const *dst_ptrs, int dr, int dc, int batch, cudaStream_t s) {
  if (aligned && sstep % (int64_t)(4 * sizeof(T)) == 0)
    run_bayer_gray_pyrdown_p<T, true>(src_ptrs, sstep, rows, cols, ry, rx, scale, dst_ptrs, dr, dc, batch, s);
  else
    run_bayer_gray_pyrdown_p<T, false>(src_ptrs, sstep, rows, cols, ry, rx, scale, dst_ptrs, dr, dc, batch, s);
}

}  // namespace

int launch_debayer_nn2(const void *src, int64_t sstep, int depth, int rows, int cols, int colorid, void *dst, int64_t dstep,
                       cudaStream_t s) {
  SSK_REQUIRE(!(rows & 1) && !(cols & 1), "debayer_nn2: Can not make debayer for uneven image size");
  int ry, rx;
  switch (colorid) {
    case SSK_COLORID_BAYER_RGGB: ry = 0; rx = 0; break;
    case SSK_COLORID_BAYER_GRBG: ry = 0; rx = 1; break;
    case SSK_COLORID_BAYER_GBRG: ry = 1; rx = 0; break;
    case SSK_COLORID_BAYER_BGGR: ry = 1; rx = 1; break;
    default: set_error("debayer_nn2: unsupported colorid (RGGB, GRBG, GBRG, BGGR)"); return SSK_ERR_INVALID;
  }
  if (depth == SSK_8U) run_debayer<uint8_t>(src, sstep, rows, cols, ry, rx, dst, dstep, s);
  else if (depth == SSK_16U) run_debayer<uint16_t>(src, sstep, rows, cols, ry, rx, dst, dstep, s);
  else if (depth == SSK_32F) run_debayer<float>(src, sstep, rows, cols, ry, rx, dst, dstep, s);
  else { set_error("debayer_nn2: CV_8U, CV_16U or CV_32F"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

// batched: src_ptrs[b] raw Bayer frames (rows x cols, one channel, `depth`; frames_aligned16: every pointer is a multiple of 16), dst_ptrs[b] dense CV_32FC1 images of
// ((rows + 1) / 2) x ((cols + 1) / 2): pyrDown(gray(debayer_nn2(raw) * scale))
int launch_bayer_gray_pyrdown(const void *const *src_ptrs, bool frames_aligned16, int64_t sstep, int depth, int rows, int cols, int colorid, float scale,
                              float *const *dst_ptrs, int batch, cudaStream_t s) {
  SSK_REQUIRE(!(rows & 1) && !(cols & 1), "debayer_nn2: Can not make debayer for uneven image size");
  SSK_REQUIRE(rows >= 4 && cols >= 4, "bayer_gray_pyrdown: frame too small");
  int ry, rx;
  switch (colorid) {
    case SSK_COLORID_BAYER_RGGB: ry = 0; rx = 0; break;
    case SSK_COLORID_BAYER_GRBG: ry = 0; rx = 1; break;
    case SSK_COLORID_BAYER_GBRG: ry = 1; rx = 0; break;
    case SSK_COLORID_BAYER_BGGR: ry = 1; rx = 1; break;
    default: set_error("debayer_nn2: unsupported colorid (RGGB, GRBG, GBRG, BGGR)"); return SSK_ERR_INVALID;
  }
  const int dr = (rows + 1) / 2, dc = (cols + 1) / 2;
  if (depth == SSK_8U) run_bayer_gray_pyrdown<uint8_t>(src_ptrs, frames_aligned16, sstep, rows, cols, ry, rx, scale, dst_ptrs, dr, dc, batch, s);
  else if (depth == SSK_16U) run_bayer_gray_pyrdown<uint16_t>(src_ptrs, frames_aligned16, sstep, rows, cols, ry, rx, scale, dst_ptrs, dr, dc, batch, s);
  else if (depth == SSK_32F) run_bayer_gray_pyrdown<float>(src_ptrs, frames_aligned16, sstep, rows, cols, ry, rx, scale, dst_ptrs, dr, dc, batch, s);
  else { set_error("bayer_gray_pyrdown: CV_8U, CV_16U or CV_32F"); return SSK_ERR_INVALID; }
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk
