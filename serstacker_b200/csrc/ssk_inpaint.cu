// average_pyramid_inpaint (core/proc/inpaint/average_pyramid_inpaint.cc:17-127): the finishing step the stacking
// pipeline applies to c_frame_accumulation::compute()'s result (c_image_stacking_pipeline.cc:763-767).
//
// The reference recurses: filter2 (3x3 un-normalised box sums of image and mask, BORDER_REPLICATE, valid pixels kept,
// holes = sum / count) -> downstrike_even (keep pixel (min(2y+1,rows-1), min(2x+1,cols-1))) -> recurse while the mask
// has holes -> upject_even (dst(2y+1,2x+1) = src(y,x), zeros elsewhere, mask 1 on the injected pixels) -> filter2 with
// the level's own image as fallback.  On the device this is two short chains of HBM-bound kernels:
//   * down: level k+1 is evaluated only on the pixels downstrike keeps (the other 3/4 of the filtered level are never
//     formed),
//   * up:   each hole gathers the odd/odd pixels of its clamped 3x3 window straight from level k+1 (the up-jected
//     image is never formed), valid pixels are left untouched in place.
// Every level down to min(cols, rows) <= 1 is visited: where the reference stops early because a level has no holes,
// the deeper levels cannot change anything (the fallback wins on every pixel), so the results are identical and no
// host round trip is needed between levels.  Box sums: fp64 sum of the window rounded once to fp32 (cv::boxFilter's
// RowSum<float,double> / ColumnSum<double,float>; oracle/inpaint.py::box_sum_model, pinned against cv2).
#include "ssk_prep.cuh"

namespace ssk {
namespace {

constexpr int kInpaintMaxLevels = 32;

struct LevelGeom { int rows, cols; size_t img_off, msk_off; };   // offsets in floats into the work buffer

int build_levels(int rows, int cols, int cn, int max_levels, LevelGeom *lv, size_t *total_floats) {
  int n = 0;
  size_t off = 0;
  int r = rows, c = cols;
  for (;;) {
    lv[n].rows = r; lv[n].cols = c;
    lv[n].img_off = off; off += (size_t)r * c * cn;
    lv[n].msk_off = off; off += (size_t)r * c;
    off = (off + 3) & ~(size_t)3;
    ++n;
    // average_pyramid_recurse: descend while min(cols, rows) > 1 && max_levels > 0
    if (!(r > 1 && c > 1) || n - 1 >= max_levels || n >= kInpaintMaxLevels) break;
    r = (r + 1) / 2; c = (c + 1) / 2;
  }
  *total_floats = off;
  return n;   // number of levels including level 0; the recursion depth is n - 1
}

__device__ __forceinline__ int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// level 0: src.copyTo(img, mask); mask.convertTo(msk, CV_32F, 1/255); counts the non-zero mask pixels
template <int CN>
__global__ void __launch_bounds__(256) k_inpaint_prepare(const float *__restrict__ src, int64_t sstep, const uint8_t *__restrict__ mask,
                                                         int64_t mstep, int rows, int cols, float *__restrict__ I,
                                                         float *__restrict__ M, int *nonzero) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  int nz = 0;
  if (x < cols && y < rows) {
    const uint8_t v = __ldg(mask + (int64_t)y * mstep + x);
    nz = v != 0;
    const float *sp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(src) + (int64_t)y * sstep) + (int64_t)x * CN;
    float *o = I + ((int64_t)y * cols + x) * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) o[c] = v ? __ldg(sp + c) : 0.f;
    M[(int64_t)y * cols + x] = __fmul_rn((float)v, (float)(1.0 / 255.0));
  }
  const int cnt = __syncthreads_count(nz);
  if (threadIdx.x == 0 && cnt) atomicAdd(nonzero, cnt);
}

// filter2(level k) sampled by downstrike_even -> level k+1
template <int CN>
__global__ void __launch_bounds__(256) k_inpaint_down(const float *__restrict__ I, const float *__restrict__ M, int rows, int cols,
                                                      float *__restrict__ Id, float *__restrict__ Md, int drows, int dcols) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dcols || y >= drows) return;
  const int sy = min(2 * y + 1, rows - 1), sx = min(2 * x + 1, cols - 1);
  const float m = __ldg(M + (int64_t)sy * cols + sx);
  float out[CN], om;
  if (m != 0.f) {
#pragma unroll
    for (int c = 0; c < CN; ++c) out[c] = __ldg(I + ((int64_t)sy * cols + sx) * CN + c);
    om = 1.f;
  } else {
    double s[CN], ms = 0.0;
#pragma unroll
    for (int c = 0; c < CN; ++c) s[c] = 0.0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = clampi(sy + dy, rows - 1);
      const int x0 = clampi(sx - 1, cols - 1), x2 = clampi(sx + 1, cols - 1);
      const float *mr = M + (int64_t)yy * cols;
      const float *ir = I + (int64_t)yy * cols * CN;
      ms = __dadd_rn(ms, __dadd_rn(__dadd_rn((double)__ldg(mr + x0), (double)__ldg(mr + sx)), (double)__ldg(mr + x2)));
#pragma unroll
      for (int c = 0; c < CN; ++c)
        s[c] = __dadd_rn(s[c], __dadd_rn(__dadd_rn((double)__ldg(ir + (int64_t)x0 * CN + c), (double)__ldg(ir + (int64_t)sx * CN + c)),
                                         (double)__ldg(ir + (int64_t)x2 * CN + c)));
    }
    const float fm = (float)ms;
    if (fm != 0.f) {
      const float scale = __fdiv_rn(1.0f, fm);
#pragma unroll
      for (int c = 0; c < CN; ++c) out[c] = __fmul_rn((float)s[c], scale);
      om = 1.f;
    } else {
#pragma unroll
      for (int c = 0; c < CN; ++c) out[c] = (float)s[c];
      om = fm;
    }
  }
#pragma unroll
  for (int c = 0; c < CN; ++c) Id[((int64_t)y * dcols + x) * CN + c] = out[c];
  Md[(int64_t)y * dcols + x] = om;
}

// filter2(upject_even(level k+1)) with level k as the fallback.  TOP: writes the final image and the 8-bit mask
// (msk.convertTo(CV_8U, 255)); otherwise the holes of level k are filled in place.
template <int CN, bool TOP>
__global__ void __launch_bounds__(256) k_inpaint_up(const float *__restrict__ L, int dcols, float *I, const float *__restrict__ M,
                                                    int rows, int cols, float *__restrict__ dst, uint8_t *__restrict__ dstmask) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int64_t p = (int64_t)y * cols + x;
  const float m = __ldg(M + p);
  float out[CN];
  bool valid = true;
  if (m != 0.f) {
    if (!TOP) return;
#pragma unroll
    for (int c = 0; c < CN; ++c) out[c] = I[p * CN + c];
  } else {
    double s[CN];
    int cnt = 0;
#pragma unroll
    for (int c = 0; c < CN; ++c) s[c] = 0.0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = clampi(y + dy, rows - 1);
      double r[CN];
#pragma unroll
      for (int c = 0; c < CN; ++c) r[c] = 0.0;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = clampi(x + dx, cols - 1);
        if ((yy & 1) && (xx & 1)) {
          const float *lp = L + ((int64_t)(yy >> 1) * dcols + (xx >> 1)) * CN;
#pragma unroll
          for (int c = 0; c < CN; ++c) r[c] = __dadd_rn(r[c], (double)__ldg(lp + c));
          ++cnt;
        }
      }
#pragma unroll
      for (int c = 0; c < CN; ++c) s[c] = __dadd_rn(s[c], r[c]);
    }
    if (cnt) {
      const float scale = __fdiv_rn(1.0f, (float)cnt);
#pragma unroll
      for (int c = 0; c < CN; ++c) out[c] = __fmul_rn((float)s[c], scale);
    } else {
#pragma unroll
      for (int c = 0; c < CN; ++c) out[c] = (float)s[c];
      valid = false;
    }
  }
  if (TOP) {
#pragma unroll
    for (int c = 0; c < CN; ++c) dst[p * CN + c] = out[c];
    if (dstmask) dstmask[p] = valid ? 255 : 0;
  } else {
#pragma unroll
    for (int c = 0; c < CN; ++c) I[p * CN + c] = out[c];
  }
}

// max_levels <= 0 or a 1-pixel-wide image: the masked copy and the round trip of the mask through float
template <int CN>
__global__ void __launch_bounds__(256) k_inpaint_passthrough(const float *__restrict__ I, const float *__restrict__ M, int64_t n,
                                                             float *__restrict__ dst, uint8_t *__restrict__ dstmask) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int c = 0; c < CN; ++c) dst[i * CN + c] = I[i * CN + c];
  if (dstmask) {
    const int v = __float2int_rn(__fmul_rn(M[i], 255.0f));
    dstmask[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

template <int CN>
int run_inpaint(const float *src, int64_t sstep, const uint8_t *mask, int64_t mstep, int rows, int cols, int max_levels,
                float *work, int *d_count, float *dst, uint8_t *dstmask, int *was_full, cudaStream_t s) {
  LevelGeom lv[kInpaintMaxLevels];
  size_t total;
  const int nlev = build_levels(rows, cols, CN, max_levels, lv, &total);
  auto grid = [](int r, int c) { return dim3(div_up(c, 32), div_up(r, 8)); };
  SSK_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
  k_inpaint_prepare<CN><<<grid(rows, cols), 256, 0, s>>>(src, sstep, mask, mstep, rows, cols, work + lv[0].img_off, work + lv[0].msk_off, d_count);
  SSK_LAUNCH_CHECK();
  // average_pyramid_inpaint.cc:104-110: a mask without holes returns copies of the inputs
  int nonzero = 0;
  SSK_CUDA(cudaMemcpyAsync(&nonzero, d_count, sizeof(int), cudaMemcpyDeviceToHost, s));
  SSK_CUDA(cudaStreamSynchronize(s));
  *was_full = nonzero == rows * cols;
  if (*was_full) {
    SSK_CUDA(cudaMemcpy2DAsync(dst, (size_t)cols * CN * 4, src, sstep, (size_t)cols * CN * 4, rows, cudaMemcpyDeviceToDevice, s));
    if (dstmask) SSK_CUDA(cudaMemcpy2DAsync(dstmask, cols, mask, mstep, cols, rows, cudaMemcpyDeviceToDevice, s));
    return SSK_OK;
  }
  if (nlev == 1) {
    const int64_t n = (int64_t)rows * cols;
    k_inpaint_passthrough<CN><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(work + lv[0].img_off, work + lv[0].msk_off, n, dst, dstmask);
    SSK_LAUNCH_CHECK();
    return SSK_OK;
  }
  for (int k = 0; k + 1 < nlev; ++k) {
    k_inpaint_down<CN><<<grid(lv[k + 1].rows, lv[k + 1].cols), 256, 0, s>>>(work + lv[k].img_off, work + lv[k].msk_off, lv[k].rows, lv[k].cols,
                                                                           work + lv[k + 1].img_off, work + lv[k + 1].msk_off,
                                                                           lv[k + 1].rows, lv[k + 1].cols);
    SSK_LAUNCH_CHECK();
  }
  for (int k = nlev - 2; k >= 1; --k) {
    k_inpaint_up<CN, false><<<grid(lv[k].rows, lv[k].cols), 256, 0, s>>>(work + lv[k + 1].img_off, lv[k + 1].cols, work + lv[k].img_off,
                                                                         work + lv[k].msk_off, lv[k].rows, lv[k].cols, nullptr, nullptr);
    SSK_LAUNCH_CHECK();
  }
  k_inpaint_up<CN, true><<<grid(rows, cols), 256, 0, s>>>(work + lv[1].img_off, lv[1].cols, work + lv[0].img_off, work + lv[0].msk_off, rows,
                                                          cols, dst, dstmask);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace

size_t inpaint_work_bytes(int rows, int cols, int cn, int max_levels) {
  LevelGeom lv[kInpaintMaxLevels];
  size_t total;
  build_levels(rows, cols, cn, max_levels, lv, &total);
  return total * sizeof(float) + 16;     // + the non-zero counter
}

int launch_average_pyramid_inpaint(const float *src, int64_t sstep, const uint8_t *mask, int64_t mstep, int rows, int cols, int cn,
                                   int max_levels, void *work, float *dst, uint8_t *dstmask, int *was_full, cudaStream_t s) {
  SSK_REQUIRE(cn >= 1 && cn <= 4, "average_pyramid_inpaint: 1 to 4 channels");
  LevelGeom lv[kInpaintMaxLevels];
  size_t total;
  build_levels(rows, cols, cn, max_levels, lv, &total);
  float *w = static_cast<float *>(work);
  int *d_count = reinterpret_cast<int *>(w + total);
  int full = 0;
  int e;
  switch (cn) {
    case 1: e = run_inpaint<1>(src, sstep, mask, mstep, rows, cols, max_levels, w, d_count, dst, dstmask, &full, s); break;
    case 2: e = run_inpaint<2>(src, sstep, mask, mstep, rows, cols, max_levels, w, d_count, dst, dstmask, &full, s); break;
    case 3: e = run_inpaint<3>(src, sstep, mask, mstep, rows, cols, max_levels, w, d_count, dst, dstmask, &full, s); break;
    default: e = run_inpaint<4>(src, sstep, mask, mstep, rows, cols, max_levels, w, d_count, dst, dstmask, &full, s); break;
  }
  if (was_full) *was_full = full;
  return e;
}

}  // namespace ssk
