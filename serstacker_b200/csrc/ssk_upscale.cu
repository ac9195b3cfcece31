// upscale_image / upscale_remap / upscale_optflow of c_image_stacking_pipeline (c_image_stacking_pipeline.cc:1869-2002) as
// stand-alone operators; the per-frame loop evaluates the same device functions (ssk_upscale.cuh) inside its fused kernel.
#include <cstring>
#include "ssk_engine.cuh"
#include "ssk_upscale.cuh"

namespace ssk {

namespace {

__global__ void __launch_bounds__(256) k_upscale_f32(const UpscaleGeom g, const float *src, int64_t sstep, int cn, float *dst, int64_t dstep,
                                                     float post_scale) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= g.dw || oy >= g.dh) return;
  float *d = reinterpret_cast<float *>(reinterpret_cast<char *>(dst) + (int64_t)oy * dstep) + (int64_t)ox * cn;
  for (int c = 0; c < cn; ++c) {
    const float v = upscale_sample(g, [&](int x, int y) {
      return __ldg(reinterpret_cast<const float *>(reinterpret_cast<const char *>(src) + (int64_t)y * sstep) + (int64_t)x * cn + c); }, ox, oy);
    d[c] = post_scale == 1.f ? v : __fmul_rn(v, post_scale);
  }
}

__global__ void __launch_bounds__(256) k_upscale_mask(const UpscaleGeom g, const uint8_t *src, int64_t sstep, uint8_t *dst, int64_t dstep) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= g.dw || oy >= g.dh) return;
  dst[(int64_t)oy * dstep + ox] = upscale_mask(g, [&](int x, int y) { return __ldg(src + (int64_t)y * sstep + x) == 255; }, ox, oy) ? 255 : 0;
}

int up_check_mat(const ssk_mat *m, const char *what) {
  if (!m || !m->data || m->rows <= 0 || m->cols <= 0) { set_error(std::string(what) + ": empty image"); return SSK_ERR_INVALID; }
  const int d = type_depth(m->type), cn = type_cn(m->type);
  if (!depth_bytes(d) || cn < 1 || cn > 4) { set_error(std::string(what) + ": unsupported type"); return SSK_ERR_INVALID; }
  if (m->step < (int64_t)m->cols * cn * depth_bytes(d)) { set_error(std::string(what) + ": step smaller than a row"); return SSK_ERR_INVALID; }
  return SSK_OK;
}

struct UpScratch {
  cudaStream_t stream = nullptr;
  DevBuf a, b;
  int init() {
    if (!stream) SSK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return SSK_OK;
  }
};
UpScratch &up_scratch() { static thread_local UpScratch s; return s; }

// src (host or device) -> up-scaled dst (host or device) through dense device scratch
template <class T, class LAUNCH>
int upscale_io(const ssk_mat *src, ssk_mat *dst, int option, const char *what, LAUNCH launch) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return SSK_ERR_CUDA;
  }
  SSK_REQUIRE(option >= SSK_UPSCALE_NONE && option <= SSK_UPSCALE_X30, "upscale: option must be none / x2.0 / x1.5 / x3.0");
  if (int e = up_check_mat(src, what)) return e;
  if (int e = up_check_mat(dst, what)) return e;
  int dw, dh;
  upscale_size(option, src->cols, src->rows, &dw, &dh);
  SSK_REQUIRE(dst->type == src->type && dst->cols == dw && dst->rows == dh, "upscale: dst must have the source type and the up-scaled size");
  SSK_REQUIRE(option != SSK_UPSCALE_PYRUP || (src->cols >= 2 && src->rows >= 2), "upscale: pyrUp needs a source of at least 2 x 2");
  UpScratch &sc = up_scratch();
  if (int e = sc.init()) return e;
  const int cn = type_cn(src->type);
  const size_t srow = (size_t)src->cols * cn * sizeof(T), drow = (size_t)dw * cn * sizeof(T);
  const T *ds; int64_t dsstep;
  if (src->mem == SSK_MEM_DEVICE) { ds = static_cast<const T *>(src->data); dsstep = src->step; }
  else {
    if (int e = sc.a.ensure(srow * src->rows)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(sc.a.p, srow, src->data, src->step, srow, src->rows, cudaMemcpyHostToDevice, sc.stream));
    ds = sc.a.as<T>(); dsstep = (int64_t)srow;
  }
  T *dd; int64_t ddstep;
  if (dst->mem == SSK_MEM_DEVICE) { dd = static_cast<T *>(dst->data); ddstep = dst->step; }
  else { if (int e = sc.b.ensure(drow * dh)) return e; dd = sc.b.as<T>(); ddstep = (int64_t)drow; }
  if (int e = launch(ds, dsstep, dd, ddstep, cn, sc.stream)) return e;
  if (dst->mem != SSK_MEM_DEVICE) SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, dd, drow, drow, dh, cudaMemcpyDeviceToHost, sc.stream));
  SSK_CUDA(cudaStreamSynchronize(sc.stream));
  return SSK_OK;
}

}  // namespace

int launch_upscale_f32(int option, const float *src, int64_t sstep, int rows, int cols, int cn, float *dst, int64_t dstep, float post_scale,
                       cudaStream_t s) {
  const UpscaleGeom g = make_upscale_geom(option, cols, rows);
  dim3 grid(div_up(g.dw, 32), div_up(g.dh, 8));
  k_upscale_f32<<<grid, 256, 0, s>>>(g, src, sstep, cn, dst, dstep, post_scale);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int launch_upscale_mask(int option, const uint8_t *src, int64_t sstep, int rows, int cols, uint8_t *dst, int64_t dstep, cudaStream_t s) {
  const UpscaleGeom g = make_upscale_geom(option, cols, rows);
  dim3 grid(div_up(g.dw, 32), div_up(g.dh, 8));
  k_upscale_mask<<<grid, 256, 0, s>>>(g, src, sstep, dst, dstep);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

}  // namespace ssk

using namespace ssk;

extern "C" {

int ssk_upscale_size(int option, int cols, int rows, int *ucols, int *urows) {
  SSK_REQUIRE(option >= SSK_UPSCALE_NONE && option <= SSK_UPSCALE_X30 && ucols && urows, "upscale: bad argument");
  upscale_size(option, cols, rows, ucols, urows);
  return SSK_OK;
}

int ssk_upscale_image(int option, const ssk_mat *src, const ssk_mat *srcmask, ssk_mat *dst, ssk_mat *dstmask) {
  if (src && dst) {
    SSK_REQUIRE(type_depth(src->type) == SSK_32F, "upscale_image: CV_32F images (frames are CV_32F when the pipeline up-scales them)");
    if (int e = upscale_io<float>(src, dst, option, "upscale_image", [&](const float *s, int64_t ss, float *d, int64_t dstp, int cn, cudaStream_t st) {
          return launch_upscale_f32(option, s, ss, src->rows, src->cols, cn, d, dstp, 1.f, st); })) return e;
  }
  if (srcmask && dstmask) {
    SSK_REQUIRE(srcmask->type == SSK_8UC1, "upscale_image: the mask must be CV_8UC1");
    if (int e = upscale_io<uint8_t>(srcmask, dstmask, option, "upscale_image mask", [&](const uint8_t *s, int64_t ss, uint8_t *d, int64_t dstp, int, cudaStream_t st) {
          return launch_upscale_mask(option, s, ss, srcmask->rows, srcmask->cols, d, dstp, st); })) return e;
  }
  return SSK_OK;
}

int ssk_upscale_remap(int option, const ssk_mat *srcmap, ssk_mat *dstmap) {
  SSK_REQUIRE(srcmap && dstmap && srcmap->type == SSK_32FC2, "upscale_remap: CV_32FC2 maps");
  return upscale_io<float>(srcmap, dstmap, option, "upscale_remap", [&](const float *s, int64_t ss, float *d, int64_t dstp, int cn, cudaStream_t st) {
    return launch_upscale_f32(option, s, ss, srcmap->rows, srcmap->cols, cn, d, dstp, 1.f, st); });
}

int ssk_upscale_optflow(int option, const ssk_mat *srcmap, ssk_mat *dstmap) {
  SSK_REQUIRE(srcmap && dstmap && srcmap->type == SSK_32FC2, "upscale_optflow: CV_32FC2 flows");
  // cv::multiply(dstmap, factor, dstmap) after the up-scaling (c_image_stacking_pipeline.cc:1918-1940)
  const float f = option == SSK_UPSCALE_PYRUP ? 2.f : option == SSK_UPSCALE_X15 ? 1.5f : option == SSK_UPSCALE_X30 ? 3.f : 1.f;
  return upscale_io<float>(srcmap, dstmap, option, "upscale_optflow", [&](const float *s, int64_t ss, float *d, int64_t dstp, int cn, cudaStream_t st) {
    return launch_upscale_f32(option, s, ss, srcmap->rows, srcmap->cols, cn, d, dstp, f, st); });
}

}  // extern "C"
