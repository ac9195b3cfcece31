// c_canvas_average on the device (core/average/c_frame_accumulation.h:65-137, c_frame_accumulation.cc:264-445): a weighted
// running mean on a canvas larger than the frames.  Every frame is remapped (cv::remap, BORDER_REPLICATE) into a bounding box
// of the canvas; when a box comes within 32 px of a canvas edge the canvas content is shifted by 64 px.
//
// The reference runs cv::remap of the frame, cv::remap of the weights (INTER_NEAREST for an 8U mask, INTER_LINEAR for CV_32F
// weights, BORDER_CONSTANT 0) and the accumulator update as three passes; here they are one kernel over the box: each thread
// samples frame and weight at its map coordinate and updates the canvas pixel in place (one read of the map, one
// read-modify-write of the canvas, gathers served by L1/L2).
#include <cmath>
#include <cstring>
#include <new>
#include "ssk_engine.cuh"
#include "ssk_fused_impl.cuh"

namespace ssk {
namespace {

struct CanvasAddArgs {
  Img src;                                 // CV_32F, cn 1..4
  const void *weights; int64_t w_step; int wtype;   // -1 none, SSK_8UC1 mask, SSK_32FC1 weights
  const float2 *rmap; int64_t rmap_step;   // bytes; null: no remap (the frame is added as it is)
  int interp;
  float *acc; int64_t acc_step;            // canvas view at the box origin; steps in floats
  float *wacc; int64_t wacc_step;
  int rows, cols;                          // box size
};

__global__ void __launch_bounds__(256) k_canvas_add(const CanvasAddArgs a, const Tables tab) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.cols || y >= a.rows) return;
  float u = (float)x, v = (float)y;
  if (a.rmap) {
    const float2 m = *reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(a.rmap) + (int64_t)y * a.rmap_step + (int64_t)x * 8);
    u = m.x; v = m.y;
  }
  float wk = 1.f;
  bool weighted = false;
  if (a.wtype == SSK_8UC1) {
    uint8_t mv;
    if (a.rmap) {                          // remap(mask, INTER_NEAREST, BORDER_CONSTANT 0)
      const int ix = __float2int_rn(u), iy = __float2int_rn(v);
      mv = ((unsigned)ix < (unsigned)a.src.cols && (unsigned)iy < (unsigned)a.src.rows)
               ? static_cast<const uint8_t *>(a.weights)[(int64_t)iy * a.w_step + ix] : 0;
    } else mv = static_cast<const uint8_t *>(a.weights)[(int64_t)y * a.w_step + x];
    if (!mv) return;
  } else if (a.wtype == SSK_32FC1) {
    weighted = true;
    if (a.rmap) {                          // remap(weights, INTER_LINEAR, BORDER_CONSTANT 0)
      Img wi; wi.data = a.weights; wi.step = a.w_step; wi.rows = a.src.rows; wi.cols = a.src.cols; wi.depth = SSK_32F; wi.cn = 1; wi.scale = 1.f;
      wk = sample_any(wi, 0, u, v, SSK_INTER_LINEAR, SSK_BORDER_CONSTANT, 0.f, tab.cubic);
    } else wk = *reinterpret_cast<const float *>(static_cast<const char *>(a.weights) + (int64_t)y * a.w_step + (int64_t)x * 4);
    if (!(wk > 0.f)) return;
  }
  float *W = a.wacc + (int64_t)y * a.wacc_step + x;
  float *A = a.acc + (int64_t)y * a.acc_step + (int64_t)x * a.src.cn;
  const float Wn = *W + wk;
  const float factor = weighted ? __fdiv_rn(wk, Wn) : __fdiv_rn(1.0f, Wn);
  *W = Wn;
  for (int c = 0; c < a.src.cn; ++c) {
    const float I = a.rmap ? sample_any(a.src, c, u, v, a.interp, SSK_BORDER_REPLICATE, 0.f, tab.cubic, tab.lanczos)
                           : *reinterpret_cast<const float *>(static_cast<const char *>(a.src.data) + (int64_t)y * a.src.step + ((int64_t)x * a.src.cn + c) * 4);
    A[c] = fmaf(I - A[c], factor, A[c]);
  }
}

// maintainCanvasBoundaries: dst = 0, dst(dst_roi) = src(src_roi)
__global__ void __launch_bounds__(256) k_canvas_shift(const float *src, float *dst, int rows, int cols, int cn, int sx, int sy) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const int px = x - sx, py = y - sy;
  const bool in = (unsigned)px < (unsigned)cols && (unsigned)py < (unsigned)rows;
  for (int c = 0; c < cn; ++c) dst[((int64_t)y * cols + x) * cn + c] = in ? src[((int64_t)py * cols + px) * cn + c] : 0.f;
}

struct Rect { int x, y, w, h; };
Rect intersect(Rect a, Rect b) {
  const int x0 = std::max(a.x, b.x), y0 = std::max(a.y, b.y), x1 = std::min(a.x + a.w, b.x + b.w), y1 = std::min(a.y + a.h, b.y + b.h);
  if (x1 > x0 && y1 > y0) return {x0, y0, x1 - x0, y1 - y0};
  return {0, 0, 0, 0};
}

int canvas_check_mat(const ssk_mat *m, const char *what) {
  if (!m || !m->data || m->rows <= 0 || m->cols <= 0) { set_error(std::string(what) + ": empty image"); return SSK_ERR_INVALID; }
  const int d = type_depth(m->type), cn = type_cn(m->type);
  if (!depth_bytes(d) || cn < 1 || cn > 4) { set_error(std::string(what) + ": unsupported type"); return SSK_ERR_INVALID; }
  if (m->step < (int64_t)m->cols * cn * depth_bytes(d)) { set_error(std::string(what) + ": step smaller than a row"); return SSK_ERR_INVALID; }
  return SSK_OK;
}

}  // namespace
}  // namespace ssk

using namespace ssk;

struct ssk_canvas {
  cudaStream_t stream = nullptr;
  int interpolation = SSK_INTER_LINEAR;
  int want_cols = 0, want_rows = 0;        // setCanvasSize
  int rows = 0, cols = 0, cn = 0;          // canvas (0: empty)
  int frames = 0;
  Rect last = {0, 0, 0, 0};
  DevBuf acc, wacc, acc2, wacc2, st_img, st_w, st_map, st_out;
  Tables tab;
  ~ssk_canvas() { if (stream) cudaStreamDestroy(stream); }
};

namespace ssk {
namespace {

// a host / device image as a dense-or-strided device view
int canvas_to_device(ssk_canvas *h, const ssk_mat *m, DevBuf &staging, const void **data, int64_t *step) {
  if (m->mem == SSK_MEM_DEVICE) { *data = m->data; *step = m->step; return SSK_OK; }
  const size_t rowb = (size_t)m->cols * type_cn(m->type) * depth_bytes(type_depth(m->type));
  if (int e = staging.ensure(rowb * m->rows)) return e;
  SSK_CUDA(cudaMemcpy2DAsync(staging.p, rowb, m->data, m->step, rowb, m->rows, cudaMemcpyHostToDevice, h->stream));
  *data = staging.p; *step = (int64_t)rowb;
  return SSK_OK;
}

// maintainCanvasBoundaries (c_frame_accumulation.cc:272-322)
int canvas_maintain(ssk_canvas *h, Rect &bbox) {
  if (!h->rows || bbox.w <= 0 || bbox.h <= 0) return SSK_OK;
  const int margin = 32;
  int sx = 0, sy = 0;
  if (bbox.x < margin) sx = 2 * margin; else if (bbox.x + bbox.w >= h->cols - margin) sx = -2 * margin;
  if (bbox.y < margin) sy = 2 * margin; else if (bbox.y + bbox.h >= h->rows - margin) sy = -2 * margin;
  if (!sx && !sy) return SSK_OK;
  if (h->cols - std::abs(sx) <= 0 || h->rows - std::abs(sy) <= 0) return SSK_OK;
  const size_t n = (size_t)h->rows * h->cols;
  if (int e = h->acc2.ensure(n * h->cn * 4)) return e;
  if (int e = h->wacc2.ensure(n * 4)) return e;
  dim3 grid(div_up(h->cols, 32), div_up(h->rows, 8));
  k_canvas_shift<<<grid, 256, 0, h->stream>>>(h->acc.as<float>(), h->acc2.as<float>(), h->rows, h->cols, h->cn, sx, sy);
  SSK_LAUNCH_CHECK();
  k_canvas_shift<<<grid, 256, 0, h->stream>>>(h->wacc.as<float>(), h->wacc2.as<float>(), h->rows, h->cols, 1, sx, sy);
  SSK_LAUNCH_CHECK();
  std::swap(h->acc.p, h->acc2.p); std::swap(h->acc.bytes, h->acc2.bytes);
  std::swap(h->wacc.p, h->wacc2.p); std::swap(h->wacc.bytes, h->wacc2.bytes);
  bbox.x += sx; bbox.y += sy;
  return SSK_OK;
}

}  // namespace
}  // namespace ssk

extern "C" {

int ssk_canvas_create(int interpolation, ssk_canvas **out) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return SSK_ERR_CUDA;
  }
  SSK_REQUIRE(out, "null argument");
  SSK_REQUIRE(interpolation == SSK_INTER_NEAREST || interpolation == SSK_INTER_LINEAR || interpolation == SSK_INTER_CUBIC || interpolation == SSK_INTER_AREA ||
              interpolation == SSK_INTER_LANCZOS4, "c_canvas_average: interpolation must be NEAREST, LINEAR, CUBIC, AREA or LANCZOS4");
  ssk_canvas *h = new (std::nothrow) ssk_canvas();
  SSK_REQUIRE(h, "out of memory");
  h->interpolation = remap_interp(interpolation);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; set_error("cudaStreamCreate failed"); return SSK_ERR_CUDA; }
  if (int e = get_tables(&h->tab)) { delete h; return e; }
  *out = h;
  return SSK_OK;
}

int ssk_canvas_destroy(ssk_canvas *h) { delete h; return SSK_OK; }

int ssk_canvas_clear(ssk_canvas *h) {
  SSK_REQUIRE(h, "null handle");
  h->rows = h->cols = h->cn = 0; h->frames = 0; h->last = {0, 0, 0, 0};
  return SSK_OK;
}

int ssk_canvas_set_canvas_size(ssk_canvas *h, int cols, int rows) {
  SSK_REQUIRE(h && cols >= 0 && rows >= 0, "c_canvas_average::setCanvasSize: bad size");
  ssk_canvas_clear(h);
  h->want_cols = cols; h->want_rows = rows;
  return SSK_OK;
}

int ssk_canvas_add(ssk_canvas *h, const ssk_mat *image, const ssk_mat *weights_or_mask, const ssk_mat *rmap, const int bbox[4]) {
  SSK_REQUIRE(h, "null handle");
  if (int e = canvas_check_mat(image, "c_canvas_average::add image")) return e;
  // the accumulator takes the image's type and _weighted_average_update wants CV_32F (c_frame_accumulation.cc:357, 43-46)
  SSK_REQUIRE(type_depth(image->type) == SSK_32F, "c_canvas_average: CV_32F frames");
  const int cn = type_cn(image->type);
  if (rmap) {
    if (int e = canvas_check_mat(rmap, "c_canvas_average::add rmap")) return e;
    SSK_REQUIRE(rmap->type == SSK_32FC2, "c_canvas_average: rmap must be CV_32FC2");
    if (h->rows) SSK_REQUIRE(rmap->cols <= h->cols && rmap->rows <= h->rows, "c_canvas_average: rmap larger than the canvas");   // :330-335
  }
  int wtype = -1;
  if (weights_or_mask && weights_or_mask->data) {
    if (int e = canvas_check_mat(weights_or_mask, "c_canvas_average::add weights")) return e;
    wtype = weights_or_mask->type;
    SSK_REQUIRE(wtype == SSK_8UC1 || wtype == SSK_32FC1, "c_canvas_average: weights must be CV_8UC1 or CV_32FC1");
    SSK_REQUIRE(weights_or_mask->rows == image->rows && weights_or_mask->cols == image->cols, "c_canvas_average: weights size differs from the image size");
  }
  CanvasAddArgs a = {};
  a.src.rows = image->rows; a.src.cols = image->cols; a.src.depth = SSK_32F; a.src.cn = cn; a.src.scale = 1.f;
  if (int e = canvas_to_device(h, image, h->st_img, &a.src.data, &a.src.step)) return e;
  a.wtype = wtype;
  if (wtype >= 0)
    if (int e = canvas_to_device(h, weights_or_mask, h->st_w, &a.weights, &a.w_step)) return e;
  a.interp = h->interpolation;
  Rect roi;
  if (!h->rows) {
    // very first frame (c_frame_accumulation.cc:346-364): centred on a fresh canvas, no remap
    const int cw = std::max(h->want_cols, 3 * image->cols / 2), chh = std::max(h->want_rows, 3 * image->rows / 2);
    const size_t n = (size_t)cw * chh;
    if (int e = h->acc.ensure(n * cn * 4)) return e;
    if (int e = h->wacc.ensure(n * 4)) return e;
    SSK_CUDA(cudaMemsetAsync(h->acc.p, 0, n * cn * 4, h->stream));
    SSK_CUDA(cudaMemsetAsync(h->wacc.p, 0, n * 4, h->stream));
    h->rows = chh; h->cols = cw; h->cn = cn;
    roi = {cw / 2 - image->cols / 2, chh / 2 - image->rows / 2, image->cols, image->rows};
    a.rows = roi.h; a.cols = roi.w;
  } else {
    SSK_REQUIRE(cn == h->cn, "c_canvas_average: channel count differs from the canvas");
    const Rect canvas = {0, 0, h->cols, h->rows};
    const bool have_bbox = bbox && bbox[2] > 0 && bbox[3] > 0;
    if (!have_bbox && !rmap) {
      roi = intersect({h->last.x, h->last.y, image->cols, image->rows}, canvas);
      a.rows = image->rows; a.cols = image->cols;
    } else {
      SSK_REQUIRE(have_bbox && rmap, "c_canvas_average::add: a remap needs both rmap and new_canvas_bbox");
      roi = intersect({bbox[0], bbox[1], bbox[2], bbox[3]}, canvas);
      SSK_REQUIRE(roi.w > 0, "c_canvas_average::add: ROI is empty");   // :378-381
      if (int e = canvas_maintain(h, roi)) return e;
      const void *md; int64_t ms;
      if (int e = canvas_to_device(h, rmap, h->st_map, &md, &ms)) return e;
      a.rmap = static_cast<const float2 *>(md); a.rmap_step = ms;
      a.rows = rmap->rows; a.cols = rmap->cols;
    }
  }
  // cv::Mat::operator()(Rect) asserts the box lies inside the matrix: the 64-px shift can push it out on a small canvas
  SSK_REQUIRE(roi.x >= 0 && roi.y >= 0 && roi.x + roi.w <= h->cols && roi.y + roi.h <= h->rows,
              "c_canvas_average::add: ROI outside the canvas after maintainCanvasBoundaries");
  // _weighted_average_update rejects a size mismatch and the reference ignores that (the frame is still counted)
  if (roi.w == a.cols && roi.h == a.rows) {
    a.acc = h->acc.as<float>() + ((int64_t)roi.y * h->cols + roi.x) * cn; a.acc_step = (int64_t)h->cols * cn;
    a.wacc = h->wacc.as<float>() + (int64_t)roi.y * h->cols + roi.x; a.wacc_step = h->cols;
    dim3 grid(div_up(a.cols, 32), div_up(a.rows, 8));
    k_canvas_add<<<grid, 256, 0, h->stream>>>(a, h->tab);
    SSK_LAUNCH_CHECK();
  }
  h->last = roi;
  ++h->frames;
  SSK_CUDA(cudaStreamSynchronize(h->stream));   // host inputs may be released by the caller
  return SSK_OK;
}

int ssk_canvas_compute(ssk_canvas *h, ssk_mat *avg, ssk_mat *mask, double dscale, const int rbbox[4]) {
  SSK_REQUIRE(h, "null handle");
  SSK_REQUIRE(h->frames >= 1, "c_canvas_average::compute: no accumulated frames");   // :410-412
  Rect box = {0, 0, h->cols, h->rows};
  if (rbbox && rbbox[2] > 0 && rbbox[3] > 0) box = intersect({rbbox[0], rbbox[1], rbbox[2], rbbox[3]}, box);
  SSK_REQUIRE(box.w > 0, "c_canvas_average::compute: empty box");
  const float *A = h->acc.as<float>() + ((int64_t)box.y * h->cols + box.x) * h->cn;
  const float *W = h->wacc.as<float>() + (int64_t)box.y * h->cols + box.x;
  if (avg) {
    if (int e = canvas_check_mat(avg, "c_canvas_average::compute avg")) return e;
    SSK_REQUIRE(avg->type == SSK_MAKETYPE(SSK_32F, h->cn) && avg->rows == box.h && avg->cols == box.w, "c_canvas_average::compute: avg must be CV_32F of the box size");
  }
  if (mask) {
    if (int e = canvas_check_mat(mask, "c_canvas_average::compute mask")) return e;
    SSK_REQUIRE(mask->type == SSK_8UC1 && mask->rows == box.h && mask->cols == box.w, "c_canvas_average::compute: mask must be CV_8UC1 of the box size");
  }
  // dense scratch outputs of the box, then strided copies into the caller's images
  const size_t n = (size_t)box.w * box.h;
  if (int e = h->st_out.ensure(n * h->cn * 4 + n)) return e;
  float *d_avg = h->st_out.as<float>();
  uint8_t *d_mask = reinterpret_cast<uint8_t *>(d_avg + n * h->cn);
  // launch_acc_compute works on dense accumulators: gather the box first
  SSK_CUDA(cudaMemcpy2DAsync(d_avg, (size_t)box.w * h->cn * 4, A, (size_t)h->cols * h->cn * 4, (size_t)box.w * h->cn * 4, box.h, cudaMemcpyDeviceToDevice, h->stream));
  if (int e = h->acc2.ensure(std::max(h->acc2.bytes, n * 4))) return e;
  float *d_w = h->acc2.as<float>();
  SSK_CUDA(cudaMemcpy2DAsync(d_w, (size_t)box.w * 4, W, (size_t)h->cols * 4, (size_t)box.w * 4, box.h, cudaMemcpyDeviceToDevice, h->stream));
  if (int e = launch_acc_compute(d_avg, d_w, box.h, box.w, h->cn, (float)dscale, d_avg, (int64_t)box.w * h->cn * 4, d_mask, box.w, h->stream)) return e;
  if (avg)
    SSK_CUDA(cudaMemcpy2DAsync(avg->data, avg->step, d_avg, (size_t)box.w * h->cn * 4, (size_t)box.w * h->cn * 4, box.h,
                               avg->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  if (mask)
    SSK_CUDA(cudaMemcpy2DAsync(mask->data, mask->step, d_mask, box.w, box.w, box.h,
                               mask->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

int ssk_canvas_accumulated_frames(const ssk_canvas *h) { return h ? h->frames : 0; }

int ssk_canvas_size(const ssk_canvas *h, int *cols, int *rows, int *channels) {
  SSK_REQUIRE(h, "null handle");
  if (cols) *cols = h->cols;
  if (rows) *rows = h->rows;
  if (channels) *channels = h->cn;
  return SSK_OK;
}

int ssk_canvas_last_bbox(const ssk_canvas *h, int bbox[4]) {
  SSK_REQUIRE(h && bbox, "null argument");
  bbox[0] = h->last.x; bbox[1] = h->last.y; bbox[2] = h->last.w; bbox[3] = h->last.h;
  return SSK_OK;
}

}  // extern "C"
