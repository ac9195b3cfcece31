// debayer_nn2 (core/io/debayer.cc:827-1195): the bilinear demosaic read_input_frame applies to raw Bayer frames
// (c_image_stacking_pipeline_base.cc:125-279).  One thread per raw pixel gathers its 3x3 neighbourhood (rows / columns
// -1 -> 1 and N -> N-2, as the reference indexes its first / last row and pixel pair) and writes the BGR triple of the
// source depth: integer depths round as (n/2 + sum) / n, CV_32F sums left to right in the reference's operand order
// (diagonals: TL + TR + BL + BR; cross: T + L + R + B).  HBM-bound: 1 sample read (+ L1/L2-served neighbours), 3 written.
#include "ssk_prep.cuh"

namespace ssk {
namespace {

template <class T> struct Wide { typedef int type; };
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

}  // namespace ssk
