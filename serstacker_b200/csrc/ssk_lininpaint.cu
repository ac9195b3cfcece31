// linear_interpolation_inpaint (core/proc/inpaint/linear_interpolation_inpaint.cc:14-368): what create_reference_frame applies
// to the generated master frame (c_image_stacking_pipeline.cc:1282-1284) and read_input_frame to frames that come with a
// missing-pixel mask (c_image_stacking_pipeline_base.cc:258-261).
//
// Reference algorithm, per round: every run of holes of a row is interpolated between its two valid neighbours
// (_interpolate_holes_h2), likewise along the columns (_interpolate_holes_v2), and each hole takes the distance-weighted mix
// of the two (_fill_holes2); rounds repeat while something was filled.  Here one round is three kernels: nearest valid
// neighbour to the left / right of every pixel (one thread per row), above / below (one thread per column), and the
// per-pixel fill, which evaluates the reference's float expressions in its operand order:
//   sv + (x - s) * kk,  kk = (ev - sv) * (1.0f / (end - start)),   dd * (h * dv + v * dh),  dd = 1.0f / (dh + dv)
// (the reference is built with -ffast-math, so its own result is defined up to FMA contraction of these expressions).
#include "ssk_prep.cuh"

namespace ssk {
namespace {

__global__ void k_lin_nearest_rows(const uint8_t *mask, int64_t mstep, int rows, int cols, int *L, int *R) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= rows) return;
  const uint8_t *m = mask + (int64_t)y * mstep;
  int *l = L + (int64_t)y * cols, *r = R + (int64_t)y * cols;
  int last = -1;
  for (int x = 0; x < cols; ++x) { if (m[x]) last = x; l[x] = last; }
  last = cols;
  for (int x = cols - 1; x >= 0; --x) { if (m[x]) last = x; r[x] = last; }
}

__global__ void k_lin_nearest_cols(const uint8_t *mask, int64_t mstep, int rows, int cols, int *U, int *D) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= cols) return;
  int last = -1;
  for (int y = 0; y < rows; ++y) { if (mask[(int64_t)y * mstep + x]) last = y; U[(int64_t)y * cols + x] = last; }
  last = rows;
  for (int y = rows - 1; y >= 0; --y) { if (mask[(int64_t)y * mstep + x]) last = y; D[(int64_t)y * cols + x] = last; }
}

// one axis of a hole: value interpolated along it and the reference's distance (0: no valid neighbour on this axis)
__device__ __forceinline__ float lin_axis(float sv, float ev, int p, int s, int e, int n, float *dist) {
  const bool hs = s >= 0, he = e < n;
  if (hs && he) {
    const float scale = __fdiv_rn(1.0f, (float)(e - s - 1));
    const float kk = __fmul_rn(__fsub_rn(ev, sv), scale);
    *dist = (float)max(p - s, e - p);
    return __fadd_rn(sv, __fmul_rn((float)(p - s), kk));
  }
  if (hs) { *dist = (float)(p - s); return sv; }
  if (he) { *dist = (float)(e - p); return ev; }
  *dist = 0.f;
  return 0.f;
}

template <int CN>
__global__ void __launch_bounds__(256) k_lin_fill(float *img, int64_t istep, uint8_t *mask, int64_t mstep, int rows, int cols, const int *L,
                                                  const int *R, const int *U, const int *D, int *filled) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  bool did = false;
  if (x < cols && y < rows && !mask[(int64_t)y * mstep + x]) {
    const int64_t p = (int64_t)y * cols + x;
    const int l = L[p], r = R[p], u = U[p], d = D[p];
    float *row = reinterpret_cast<float *>(reinterpret_cast<char *>(img) + (int64_t)y * istep);
    const float *rl = row + (int64_t)max(l, 0) * CN, *rr = row + (int64_t)min(r, cols - 1) * CN;
    const float *cu = reinterpret_cast<const float *>(reinterpret_cast<const char *>(img) + (int64_t)max(u, 0) * istep) + (int64_t)x * CN;
    const float *cd = reinterpret_cast<const float *>(reinterpret_cast<const char *>(img) + (int64_t)min(d, rows - 1) * istep) + (int64_t)x * CN;
    float out[CN];
    float dh = 0.f, dv = 0.f;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
      const float h = lin_axis(rl[c], rr[c], x, l, r, cols, &dh);
      const float v = lin_axis(cu[c], cd[c], y, u, d, rows, &dv);
      if (dh > 0.f && dv > 0.f) {
        const float dd = __fdiv_rn(1.0f, __fadd_rn(dh, dv));
        out[c] = __fmul_rn(dd, __fadd_rn(__fmul_rn(h, dv), __fmul_rn(v, dh)));
      } else {
        out[c] = dv > 0.f ? v : h;
      }
    }
    if (dh > 0.f || dv > 0.f) {
#pragma unroll
      for (int c = 0; c < CN; ++c) row[(int64_t)x * CN + c] = out[c];
      mask[(int64_t)y * mstep + x] = 255;
      did = true;
    }
  }
  const int n = __syncthreads_count(did);
  if (threadIdx.x == 0 && n) atomicAdd(filled, n);
}

}  // namespace

size_t lin_inpaint_work_bytes(int rows, int cols) { return (size_t)rows * cols * 4 * sizeof(int) + 16; }

// img (CV_32F, cn channels) and mask (CV_8UC1) are updated in place on the device; work: lin_inpaint_work_bytes() bytes.
// Valid pixels are never written, and a round reads valid pixels only, so filling in place equals the reference's
// separate inpaint_h / inpaint_v images.  One host synchronisation per round (the reference's `filled < 1` test).
int launch_linear_interpolation_inpaint(float *img, int64_t istep, uint8_t *mask, int64_t mstep, int rows, int cols, int cn, void *work,
                                        cudaStream_t s) {
  SSK_REQUIRE(cn >= 1 && cn <= 4, "linear_interpolation_inpaint: 1 to 4 channels");
  const size_t n = (size_t)rows * cols;
  int *L = static_cast<int *>(work), *R = L + n, *U = R + n, *D = U + n, *d_filled = D + n;
  const dim3 grid(div_up(cols, 32), div_up(rows, 8));
  for (int round = 0; round < rows + cols; ++round) {
    SSK_CUDA(cudaMemsetAsync(d_filled, 0, sizeof(int), s));
    k_lin_nearest_rows<<<div_up(rows, 64), 64, 0, s>>>(mask, mstep, rows, cols, L, R);
    SSK_LAUNCH_CHECK();
    k_lin_nearest_cols<<<div_up(cols, 64), 64, 0, s>>>(mask, mstep, rows, cols, U, D);
    SSK_LAUNCH_CHECK();
    switch (cn) {
      case 1: k_lin_fill<1><<<grid, 256, 0, s>>>(img, istep, mask, mstep, rows, cols, L, R, U, D, d_filled); break;
      case 2: k_lin_fill<2><<<grid, 256, 0, s>>>(img, istep, mask, mstep, rows, cols, L, R, U, D, d_filled); break;
      case 3: k_lin_fill<3><<<grid, 256, 0, s>>>(img, istep, mask, mstep, rows, cols, L, R, U, D, d_filled); break;
      default: k_lin_fill<4><<<grid, 256, 0, s>>>(img, istep, mask, mstep, rows, cols, L, R, U, D, d_filled); break;
    }
    SSK_LAUNCH_CHECK();
    int filled = 0;
    SSK_CUDA(cudaMemcpyAsync(&filled, d_filled, sizeof(int), cudaMemcpyDeviceToHost, s));
    SSK_CUDA(cudaStreamSynchronize(s));
    if (filled < 1) break;
  }
  return SSK_OK;
}

}  // namespace ssk
