// Input side of the per-frame loop and the steps around the master frame:
//   ssk_ser_*                          c_ser_reader (core/io/c_ser_file.cc:272-531): the SER container the pipelines read
//   ssk_input_calibrate                read_input_frame's dark / flat correction (c_image_stacking_pipeline_base.cc:143-184)
//   ssk_average_bayer_planes           average_bayer_planes (core/io/debayer.cc:277-376), the gray proxy of a raw Bayer frame
//                                      select_master_frame ranks (c_image_stacking_pipeline_base.cc:370-378)
//   ssk_color_transform                cv::transform(image, image, color_matrix) (c_image_stacking_pipeline_base.cc:263-266)
//   ssk_linear_interpolation_inpaint   linear_interpolation_inpaint (core/proc/inpaint/linear_interpolation_inpaint.cc)
//   ssk_median_filter_bad_pixels       median_filter_bad_pixels (core/proc/bad_pixels.cc:14-70) and bayer_denoise
//   ssk_bayer_denoise                  (core/io/debayer.cc:1471-1611): both branches of read_input_frame's filter_bad_pixels
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "ssk_engine.cuh"

using namespace ssk;

namespace {

int check_mat_in(const ssk_mat *m, const char *what) {
  if (!m || !m->data || m->rows <= 0 || m->cols <= 0) { set_error(std::string(what) + ": empty image"); return SSK_ERR_INVALID; }
  const int d = type_depth(m->type), cn = type_cn(m->type);
  if (!depth_bytes(d) || cn < 1 || cn > 4) { set_error(std::string(what) + ": unsupported type"); return SSK_ERR_INVALID; }
  if (m->step < (int64_t)m->cols * cn * depth_bytes(d)) { set_error(std::string(what) + ": step smaller than a row"); return SSK_ERR_INVALID; }
  return SSK_OK;
}

int need_device() {
  if (int e = chain_drain()) return e;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return SSK_ERR_CUDA;
  }
  return SSK_OK;
}

struct InScratch {
  cudaStream_t stream = nullptr;
  DevBuf a, b, c, d, w;
  int init() {
    if (!stream) SSK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return SSK_OK;
  }
};
InScratch &in_scratch() { static thread_local InScratch s; return s; }

int upload(const ssk_mat *m, DevBuf &st, cudaStream_t s, const void **data, int64_t *step) {
  const size_t rowb = (size_t)m->cols * type_cn(m->type) * depth_bytes(type_depth(m->type));
  if (m->mem == SSK_MEM_DEVICE) { *data = m->data; *step = m->step; return SSK_OK; }
  if (int e = st.ensure(rowb * m->rows)) return e;
  SSK_CUDA(cudaMemcpy2DAsync(st.p, rowb, m->data, m->step, rowb, m->rows, cudaMemcpyHostToDevice, s));
  *data = st.p; *step = (int64_t)rowb;
  return SSK_OK;
}

// read_input_frame: convertTo(CV_32F, 1 / (1 << bpp)) for integer frames, cv::subtract(frame, dark), cv::divide(frame, flat)
// (c_image_stacking_pipeline_base.cc:143-184).  cv::divide on CV_32F: dst = src1 / src2 with 0 where src2 == 0.
template <class T>
__global__ void __launch_bounds__(256) k_calibrate(const T *src, int64_t sstep, float scale, const float *dark, int64_t dstep_, const float *flat,
                                                   int64_t fstep, int rows, int n /*cols * cn*/, float *dst, int64_t ostep) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= n || y >= rows) return;
  const T raw = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)y * sstep)[x];
  float v = sizeof(T) == 4 ? (float)raw : __fmul_rn((float)raw, scale);
  if (dark) v = __fsub_rn(v, reinterpret_cast<const float *>(reinterpret_cast<const char *>(dark) + (int64_t)y * dstep_)[x]);
  if (flat) {
    const float f = reinterpret_cast<const float *>(reinterpret_cast<const char *>(flat) + (int64_t)y * fstep)[x];
    v = f != 0.f ? __fdiv_rn(v, f) : 0.f;
  }
  reinterpret_cast<float *>(reinterpret_cast<char *>(dst) + (int64_t)y * ostep)[x] = v;
}

// average_bayer_planes, raw (single channel) form: one output pixel per 2x2 cell, (c1 + s00 + s01 + s10 + s11) / 4 with
// c1 = 2 and integer division for integer samples, float arithmetic in that order otherwise (debayer.cc:300-328)
template <class T>
__global__ void __launch_bounds__(256) k_average_bayer_planes(const T *src, int64_t sstep, int drows, int dcols, T *dst, int64_t dstep) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dcols || y >= drows) return;
  const T *s0 = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)(2 * y) * sstep) + 2 * x;
  const T *s1 = reinterpret_cast<const T *>(reinterpret_cast<const char *>(src) + (int64_t)(2 * y + 1) * sstep) + 2 * x;
  T *o = reinterpret_cast<T *>(reinterpret_cast<char *>(dst) + (int64_t)y * dstep) + x;
  if (sizeof(T) == 4) {
    *o = (T)__fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(0.f, (float)s0[0]), (float)s0[1]), (float)s1[0]), (float)s1[1]), 4.0f);
  } else {
    *o = (T)((2 + (int)s0[0] + (int)s0[1] + (int)s1[0] + (int)s1[1]) / 4);
  }
}

// cv::transform on CV_32FC3 with a 3x3 (or 3x4) float matrix: dst_c = m[c][0] * s0 + m[c][1] * s1 + m[c][2] * s2 (+ m[c][3]),
// products and sums in float in that order
__global__ void __launch_bounds__(256) k_color_transform(const float *src, int64_t sstep, int rows, int cols, const float m0, const float m1,
                                                         const float m2, const float m3, const float m4, const float m5, const float m6,
                                                         const float m7, const float m8, const float m9, const float m10, const float m11,
                                                         float *dst, int64_t dstep) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  const float *s = reinterpret_cast<const float *>(reinterpret_cast<const char *>(src) + (int64_t)y * sstep) + 3 * x;
  float *d = reinterpret_cast<float *>(reinterpret_cast<char *>(dst) + (int64_t)y * dstep) + 3 * x;
  const float a = s[0], b = s[1], c = s[2];
  d[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m0, a), __fmul_rn(m1, b)), __fmul_rn(m2, c)), m3);
  d[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m4, a), __fmul_rn(m5, b)), __fmul_rn(m6, c)), m7);
  d[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m8, a), __fmul_rn(m9, b)), __fmul_rn(m10, c)), m11);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// SER reader (host side; c_ser_reader, core/io/c_ser_file.cc:272-531, header layout c_ser_file.h:42-56)
// ------------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct ser_file_header {
  char file_id[14];
  int32_t luid;
  int32_t color_id;
  int32_t is_little_endian;
  int32_t image_width;
  int32_t image_height;
  int32_t bits_per_plane;
  int32_t frames_count;
  char observer[40];
  char instrument[40];
  char telescope[40];
  uint64_t date_time;
  uint64_t date_time_utc;
};
#pragma pack(pop)
static_assert(sizeof(ser_file_header) == 178, "SER header is 178 bytes");

struct ssk_ser {
  FILE *f = nullptr;
  ser_file_header h;
  int cn = 1, depth = SSK_8U, bytes_per_sample = 1;
  bool swap = false;                    // frame data is big-endian (this host is little-endian)
  std::vector<uint64_t> timestamps;
  ~ssk_ser() { if (f) fclose(f); }
  int64_t frame_size() const { return (int64_t)h.image_width * h.image_height * cn * bytes_per_sample; }
};

// ---- median_filter_bad_pixels (bad_pixels.cc:14-56) and bayer_denoise (debayer.cc:1471-1611) ---------------------------
// median = cv::medianBlur(image, K) (BORDER_REPLICATE), mad = cv::boxFilter(|image - median|, K x K, normalised,
// BORDER_REFLECT_101), then image = median where |median - image| > k * mad + mv (mv = 1 for integer depths, 1 / 256 for float).
// K = 5 over the interleaved channels of a mono / colour frame; K = 3 over the four half-size colour planes of a raw Bayer
// mosaic, which are addressed in place (plane c = rows 2y + (c >> 1), columns 2x + (c & 1)): the reference's
// _extract_bayer_planes copy and the write-back loop disappear.  Samples are handled as float: exact for 8 / 16-bit integers.
namespace {

// a view of `cn` planes of rows x cols samples inside one frame buffer
struct BpView {
  int64_t row_step;      // bytes between plane rows
  int64_t chan_step[4];  // byte offset of plane c
  int pix_step;          // elements between plane columns
  int rows, cols, cn;
};

template <class T> __device__ __forceinline__ T *bp_ptr(T *img, const BpView &v, int y, int x, int c) {
  return reinterpret_cast<T *>((char *)(img) + (int64_t)y * v.row_step + v.chan_step[c]) + (int64_t)x * v.pix_step;
}

// median of N (9 or 25) by forgetful selection: keep N / 2 + 2 candidates, drop the extremes, take in the next sample
template <int N> __device__ __forceinline__ float median_of(const float (&in)[N]) {
  constexpr int M = N / 2 + 2;
  float v[M];
#pragma unroll
  for (int i = 0; i < M; ++i) v[i] = in[i];
#pragma unroll
  for (int n = M; n >= 3; --n) {
#pragma unroll
    for (int i = 1; i < n; ++i) { const float lo = fminf(v[0], v[i]), hi = fmaxf(v[0], v[i]); v[0] = lo; v[i] = hi; }
#pragma unroll
    for (int i = 1; i < n - 1; ++i) { const float lo = fminf(v[i], v[n - 1]), hi = fmaxf(v[i], v[n - 1]); v[i] = lo; v[n - 1] = hi; }
    // v[0] is the minimum, v[n - 1] the maximum of the n candidates: both leave, the next sample takes slot 0
    if (n > 3) v[0] = in[N - (n - 3)];
  }
  return v[1];
}

template <class T, int K>
__global__ void __launch_bounds__(256) k_bp_median(const T *img, const BpView v, float *med, float *adiff) {
  const int xc = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (xc >= v.cols * v.cn || y >= v.rows) return;
  const int x = xc / v.cn, c = xc - x * v.cn;
  float w[K * K];
#pragma unroll
  for (int dy = 0; dy < K; ++dy) {
    const int yy = min(max(y + dy - K / 2, 0), v.rows - 1);
#pragma unroll
    for (int dx = 0; dx < K; ++dx) w[dy * K + dx] = (float)*bp_ptr(img, v, yy, min(max(x + dx - K / 2, 0), v.cols - 1), c);
  }
  const float m = median_of<K * K>(w), p = w[(K * K) / 2];
  const int64_t o = (int64_t)y * v.cols * v.cn + xc;
  med[o] = m;
  adiff[o] = fabsf(p - m);                 // cv::absdiff (exact for the integer depths, no saturation can occur)
}

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

template <class T, int K>
__global__ void __launch_bounds__(256) k_bp_replace(T *img, const BpView v, const float *med, const float *adiff, float k, float mv, int integer_depth) {
  const int xc = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (xc >= v.cols * v.cn || y >= v.rows) return;
  const int x = xc / v.cn, c = xc - x * v.cn;
  // cv::boxFilter: integer depths sum in int and round sum / K^2 to the depth (no ties: K^2 is odd); CV_32F sums in double
  double sum = 0;
  for (int dy = -(K / 2); dy <= K / 2; ++dy) {
    const int yy = reflect101(y + dy, v.rows);
    for (int dx = -(K / 2); dx <= K / 2; ++dx) sum += (double)adiff[((int64_t)yy * v.cols + reflect101(x + dx, v.cols)) * v.cn + c];
  }
  float mad;
  if (integer_depth) mad = (float)__double2int_rn(sum * (1.0 / (K * K)));
  else mad = (float)(sum * (1.0 / (K * K)));
  const int64_t o = (int64_t)y * v.cols * v.cn + xc;
  T *p = bp_ptr(img, v, y, x, c);
  const float pv = (float)*p, m = med[o];
  if (fabsf(m - pv) > __fadd_rn(__fmul_rn(k, mad), mv)) *p = (T)m;
}

template <class T, int K>
int bp_run(void *dimg, const BpView &v, float *med, float *adiff, float k, float mv, int integer_depth, cudaStream_t s) {
  const dim3 grid(div_up(v.cols * v.cn, 32), div_up(v.rows, 8));
  k_bp_median<T, K><<<grid, 256, 0, s>>>(static_cast<const T *>(dimg), v, med, adiff);
  SSK_LAUNCH_CHECK();
  k_bp_replace<T, K><<<grid, 256, 0, s>>>(static_cast<T *>(dimg), v, med, adiff, k, mv, integer_depth);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int bp_filter(ssk_mat *image, double variation_threshold, bool bayer, const char *what) {
  if (int e = need_device()) return e;
  if (int e = check_mat_in(image, what)) return e;
  const int d = type_depth(image->type), cn = type_cn(image->type);
  if (bayer) {
    SSK_REQUIRE(cn == 1, "bayer_denoise: a raw Bayer frame has one channel");
    SSK_REQUIRE(!(image->rows & 1) && !(image->cols & 1) && image->rows >= 4 && image->cols >= 4, "bayer_denoise: uneven or too small image size");
  } else SSK_REQUIRE(image->rows >= 2 && image->cols >= 2, "median_filter_bad_pixels: image smaller than 2 x 2");
  InScratch &sc = in_scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  const size_t es = depth_bytes(d), rowb = (size_t)image->cols * cn * es, n = (size_t)image->rows * image->cols * cn;
  void *dimg; int64_t dstep;
  if (image->mem == SSK_MEM_DEVICE) { dimg = image->data; dstep = image->step; }
  else {
    if (int e = sc.a.ensure(rowb * image->rows)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(sc.a.p, rowb, image->data, image->step, rowb, image->rows, cudaMemcpyHostToDevice, s));
    dimg = sc.a.p; dstep = (int64_t)rowb;
  }
  if (int e = sc.b.ensure(n * 4)) return e;
  if (int e = sc.c.ensure(n * 4)) return e;
  BpView v{};
  if (bayer) {
    v.row_step = 2 * dstep; v.pix_step = 2; v.rows = image->rows / 2; v.cols = image->cols / 2; v.cn = 4;
    for (int c = 0; c < 4; ++c) v.chan_step[c] = (c >> 1) * dstep + (c & 1) * (int64_t)es;
  } else {
    v.row_step = dstep; v.pix_step = cn; v.rows = image->rows; v.cols = image->cols; v.cn = cn;
    for (int c = 0; c < 4; ++c) v.chan_step[c] = c * (int64_t)es;
  }
  const float k = (float)variation_threshold, mv = d == SSK_32F ? 1.f / 256.f : 1.f;
  float *med = sc.b.as<float>(), *adiff = sc.c.as<float>();
  int e;
  if (d == SSK_8U) e = bayer ? bp_run<uint8_t, 3>(dimg, v, med, adiff, k, mv, 1, s) : bp_run<uint8_t, 5>(dimg, v, med, adiff, k, mv, 1, s);
  else if (d == SSK_16U) e = bayer ? bp_run<uint16_t, 3>(dimg, v, med, adiff, k, mv, 1, s) : bp_run<uint16_t, 5>(dimg, v, med, adiff, k, mv, 1, s);
  else if (d == SSK_32F) e = bayer ? bp_run<float, 3>(dimg, v, med, adiff, k, mv, 0, s) : bp_run<float, 5>(dimg, v, med, adiff, k, mv, 0, s);
  else { set_error(std::string(what) + ": unsupported depth"); return SSK_ERR_INVALID; }
  if (e) return e;
  if (image->mem != SSK_MEM_DEVICE) SSK_CUDA(cudaMemcpy2DAsync(image->data, image->step, dimg, rowb, rowb, image->rows, cudaMemcpyDeviceToHost, s));
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

}  // namespace

extern "C" {

int ssk_ser_open(const char *path, ssk_ser **out) {
  SSK_REQUIRE(path && out, "ssk_ser_open: null argument");
  ssk_ser *s = new (std::nothrow) ssk_ser();
  SSK_REQUIRE(s, "out of memory");
  s->f = fopen(path, "rb");
  if (!s->f) { delete s; set_error(std::string("ssk_ser_open: cannot open ") + path); return SSK_ERR_INVALID; }
  if (fread(&s->h, 1, sizeof(s->h), s->f) != sizeof(s->h) || strncmp(s->h.file_id, "LUCAM-RECORDER", 14) != 0) {
    delete s;
    set_error(std::string("ssk_ser_open: not a SER file: ") + path);
    return SSK_ERR_INVALID;
  }
  // "There is well-known bug with endiannes in ser file format": the flag is stored inverted (c_ser_file.cc:305)
  s->h.is_little_endian = s->h.is_little_endian == 0;
  const int bpp = s->h.bits_per_plane;
  // c_ser_file.h:17-21: -32 -> CV_32F (the author's extension), 1..8 -> CV_8U, 9..16 -> CV_16U
  if (bpp == -32) { s->depth = SSK_32F; s->bytes_per_sample = 4; }
  else if (bpp > 0 && bpp <= 8) { s->depth = SSK_8U; s->bytes_per_sample = 1; }
  else if (bpp > 8 && bpp <= 16) { s->depth = SSK_16U; s->bytes_per_sample = 2; }
  else { delete s; set_error("ssk_ser_open: unsupported bits_per_plane (8U / 16U / 32F frames)"); return SSK_ERR_INVALID; }
  s->cn = (s->h.color_id == 100 || s->h.color_id == 101) ? 3 : 1;     // COLORID_RGB / COLORID_BGR
  if (s->h.image_width < 1 || s->h.image_height < 1 || s->h.frames_count < 0) { delete s; set_error("ssk_ser_open: invalid image size"); return SSK_ERR_INVALID; }
  s->swap = !s->h.is_little_endian && s->bytes_per_sample > 1;
  // optional trailer of frames_count uint64 time stamps (c_ser_file.cc:356-397)
  fseeko(s->f, 0, SEEK_END);
  const int64_t fsize = ftello(s->f);
  const int64_t ts_off = (int64_t)sizeof(s->h) + (int64_t)s->h.frames_count * s->frame_size();
  if (fsize >= ts_off + (int64_t)s->h.frames_count * 8 && s->h.frames_count > 0) {
    s->timestamps.resize(s->h.frames_count);
    fseeko(s->f, ts_off, SEEK_SET);
    if (fread(s->timestamps.data(), 8, s->timestamps.size(), s->f) != s->timestamps.size()) s->timestamps.clear();
    else if (!s->h.is_little_endian)
      for (auto &t : s->timestamps) t = __builtin_bswap64(t);
  }
  *out = s;
  return SSK_OK;
}

int ssk_ser_close(ssk_ser *s) { delete s; return SSK_OK; }

int ssk_ser_info(const ssk_ser *s, int *cols, int *rows, int *type, int *bits_per_plane, int *color_id, int *frames, int *has_timestamps) {
  SSK_REQUIRE(s, "null handle");
  if (cols) *cols = s->h.image_width;
  if (rows) *rows = s->h.image_height;
  if (type) *type = SSK_MAKETYPE(s->depth, s->cn);
  if (bits_per_plane) *bits_per_plane = s->h.bits_per_plane;
  if (color_id) *color_id = s->h.color_id;
  if (frames) *frames = s->h.frames_count;
  if (has_timestamps) *has_timestamps = s->timestamps.empty() ? 0 : 1;
  return SSK_OK;
}

int ssk_ser_read(ssk_ser *s, int frame_index, ssk_mat *dst, uint64_t *timestamp) {
  SSK_REQUIRE(s && dst && dst->data, "ssk_ser_read: null argument");
  SSK_REQUIRE(frame_index >= 0 && frame_index < s->h.frames_count, "ssk_ser_read: frame index out of range");
  SSK_REQUIRE(dst->mem == SSK_MEM_HOST && dst->type == SSK_MAKETYPE(s->depth, s->cn) && dst->rows == s->h.image_height &&
              dst->cols == s->h.image_width, "ssk_ser_read: dst must be a host image of the file's size and type");
  const int64_t rowb = (int64_t)s->h.image_width * s->cn * s->bytes_per_sample;
  SSK_REQUIRE(dst->step >= rowb, "ssk_ser_read: dst step smaller than a row");
  if (fseeko(s->f, (int64_t)sizeof(s->h) + (int64_t)frame_index * s->frame_size(), SEEK_SET) != 0) { set_error("ssk_ser_read: seek failed"); return SSK_ERR_INVALID; }
  for (int y = 0; y < s->h.image_height; ++y) {
    char *row = static_cast<char *>(dst->data) + (int64_t)y * dst->step;
    if (fread(row, 1, rowb, s->f) != (size_t)rowb) { set_error("ssk_ser_read: short read"); return SSK_ERR_INVALID; }
    if (s->swap) {
      if (s->bytes_per_sample == 2) { uint16_t *p = reinterpret_cast<uint16_t *>(row); for (int64_t i = 0; i < rowb / 2; ++i) p[i] = __builtin_bswap16(p[i]); }
      else { uint32_t *p = reinterpret_cast<uint32_t *>(row); for (int64_t i = 0; i < rowb / 4; ++i) p[i] = __builtin_bswap32(p[i]); }
    }
  }
  if (timestamp) *timestamp = s->timestamps.empty() ? 0 : s->timestamps[frame_index];
  return SSK_OK;
}

// ------------------------------------------------------------------------------------------------
int ssk_input_calibrate(const ssk_mat *frame, int bpp, const ssk_mat *dark, const ssk_mat *flat, ssk_mat *dst) {
  if (int e = need_device()) return e;
  if (int e = check_mat_in(frame, "input_calibrate frame")) return e;
  if (int e = check_mat_in(dst, "input_calibrate dst")) return e;
  const int cn = type_cn(frame->type), d = type_depth(frame->type);
  SSK_REQUIRE(dst->type == SSK_MAKETYPE(SSK_32F, cn) && dst->rows == frame->rows && dst->cols == frame->cols,
              "input_calibrate: dst must be CV_32F of the frame's size and channel count");
  for (const ssk_mat *m : {dark, flat}) {
    if (!m) continue;
    if (int e = check_mat_in(m, "input_calibrate dark/flat")) return e;
    // c_image_stacking_pipeline_base.cc:145-150, 166-171: size and channels must match the frame
    SSK_REQUIRE(m->type == SSK_MAKETYPE(SSK_32F, cn) && m->rows == frame->rows && m->cols == frame->cols,
                "darkbayer / flatbayer and input frame not match (CV_32F of the frame's size and channels)");
  }
  InScratch &sc = in_scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  const void *fp, *dp = nullptr, *lp = nullptr;
  int64_t fs, ds = 0, ls = 0;
  if (int e = upload(frame, sc.a, s, &fp, &fs)) return e;
  if (dark) if (int e = upload(dark, sc.b, s, &dp, &ds)) return e;
  if (flat) if (int e = upload(flat, sc.c, s, &lp, &ls)) return e;
  const int n = frame->cols * cn;
  const size_t orow = (size_t)n * 4;
  float *out = dst->mem == SSK_MEM_DEVICE ? static_cast<float *>(dst->data) : nullptr;
  int64_t ostep = dst->step;
  if (!out) { if (int e = sc.d.ensure(orow * frame->rows)) return e; out = sc.d.as<float>(); ostep = (int64_t)orow; }
  const dim3 grid(div_up(n, 256), frame->rows);
  const float scale = bpp_scale(d, bpp);
  if (d == SSK_32F) k_calibrate<float><<<grid, 256, 0, s>>>(static_cast<const float *>(fp), fs, 1.f, static_cast<const float *>(dp), ds, static_cast<const float *>(lp), ls, frame->rows, n, out, ostep);
  else if (d == SSK_16U) k_calibrate<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t *>(fp), fs, scale, static_cast<const float *>(dp), ds, static_cast<const float *>(lp), ls, frame->rows, n, out, ostep);
  else k_calibrate<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(fp), fs, scale, static_cast<const float *>(dp), ds, static_cast<const float *>(lp), ls, frame->rows, n, out, ostep);
  SSK_LAUNCH_CHECK();
  if (dst->mem != SSK_MEM_DEVICE)
    SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, out, orow, orow, frame->rows, cudaMemcpyDeviceToHost, s));
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

int ssk_average_bayer_planes(const ssk_mat *src, ssk_mat *dst) {
  if (int e = need_device()) return e;
  if (int e = check_mat_in(src, "average_bayer_planes src")) return e;
  if (int e = check_mat_in(dst, "average_bayer_planes dst")) return e;
  SSK_REQUIRE(type_cn(src->type) == 1, "average_bayer_planes: raw (single-channel) Bayer frames");
  SSK_REQUIRE(!(src->rows & 1) && !(src->cols & 1), "Can not average raw bayer planes for uneven image size");
  SSK_REQUIRE(dst->type == src->type && dst->rows == src->rows / 2 && dst->cols == src->cols / 2,
              "average_bayer_planes: dst is half the size, same type");
  InScratch &sc = in_scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  const void *sp; int64_t ss;
  if (int e = upload(src, sc.a, s, &sp, &ss)) return e;
  const int d = type_depth(src->type), es = depth_bytes(d);
  const size_t orow = (size_t)dst->cols * es;
  void *out = dst->mem == SSK_MEM_DEVICE ? dst->data : nullptr;
  int64_t ostep = dst->step;
  if (!out) { if (int e = sc.d.ensure(orow * dst->rows)) return e; out = sc.d.p; ostep = (int64_t)orow; }
  const dim3 grid(div_up(dst->cols, 32), div_up(dst->rows, 8));
  if (d == SSK_32F) k_average_bayer_planes<float><<<grid, 256, 0, s>>>(static_cast<const float *>(sp), ss, dst->rows, dst->cols, static_cast<float *>(out), ostep);
  else if (d == SSK_16U) k_average_bayer_planes<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t *>(sp), ss, dst->rows, dst->cols, static_cast<uint16_t *>(out), ostep);
  else k_average_bayer_planes<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(sp), ss, dst->rows, dst->cols, static_cast<uint8_t *>(out), ostep);
  SSK_LAUNCH_CHECK();
  if (dst->mem != SSK_MEM_DEVICE) SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, out, orow, orow, dst->rows, cudaMemcpyDeviceToHost, s));
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

int ssk_color_transform(const ssk_mat *src, const float *m, int mcols, ssk_mat *dst) {
  if (int e = need_device()) return e;
  if (int e = check_mat_in(src, "color_transform src")) return e;
  if (int e = check_mat_in(dst, "color_transform dst")) return e;
  SSK_REQUIRE(m && (mcols == 3 || mcols == 4), "color_transform: 3x3 or 3x4 matrix");
  SSK_REQUIRE(src->type == SSK_32FC3 && dst->type == SSK_32FC3 && dst->rows == src->rows && dst->cols == src->cols,
              "color_transform: CV_32FC3 source and destination of the same size");
  InScratch &sc = in_scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  const void *sp; int64_t ss;
  if (int e = upload(src, sc.a, s, &sp, &ss)) return e;
  const size_t orow = (size_t)dst->cols * 12;
  float *out = dst->mem == SSK_MEM_DEVICE ? static_cast<float *>(dst->data) : nullptr;
  int64_t ostep = dst->step;
  if (!out) { if (int e = sc.d.ensure(orow * dst->rows)) return e; out = sc.d.as<float>(); ostep = (int64_t)orow; }
  float k[12];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) k[r * 4 + c] = c < mcols ? m[r * mcols + c] : 0.f;
  const dim3 grid(div_up(src->cols, 32), div_up(src->rows, 8));
  k_color_transform<<<grid, 256, 0, s>>>(static_cast<const float *>(sp), ss, src->rows, src->cols, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7],
                                         k[8], k[9], k[10], k[11], out, ostep);
  SSK_LAUNCH_CHECK();
  if (dst->mem != SSK_MEM_DEVICE) SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, out, orow, orow, dst->rows, cudaMemcpyDeviceToHost, s));
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

int ssk_linear_interpolation_inpaint(const ssk_mat *src, const ssk_mat *mask, ssk_mat *dst) {
  if (int e = need_device()) return e;
  if (int e = check_mat_in(src, "linear_interpolation_inpaint src")) return e;
  if (int e = check_mat_in(dst, "linear_interpolation_inpaint dst")) return e;
  SSK_REQUIRE(type_depth(src->type) == SSK_32F && dst->type == src->type && dst->rows == src->rows && dst->cols == src->cols,
              "linear_interpolation_inpaint: CV_32F source and destination of the same size and type");
  if (mask) {
    if (int e = check_mat_in(mask, "linear_interpolation_inpaint mask")) return e;
    SSK_REQUIRE(mask->type == SSK_8UC1 && mask->rows == src->rows && mask->cols == src->cols,
                "linear_interpolation_inpaint: the mask must be CV_8UC1 of the image size");
  }
  InScratch &sc = in_scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  const int cn = type_cn(src->type);
  const size_t rowb = (size_t)src->cols * cn * 4, npx = (size_t)src->rows * src->cols;
  // work on a dense device copy (the reference copies src to dst first, :363-367)
  if (int e = sc.a.ensure(rowb * src->rows)) return e;
  SSK_CUDA(cudaMemcpy2DAsync(sc.a.p, rowb, src->data, src->step, rowb, src->rows,
                             src->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  if (mask) {
    if (int e = sc.b.ensure(npx)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(sc.b.p, src->cols, mask->data, mask->step, src->cols, src->rows,
                               mask->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    if (int e = sc.w.ensure(lin_inpaint_work_bytes(src->rows, src->cols))) return e;
    if (int e = launch_linear_interpolation_inpaint(sc.a.as<float>(), (int64_t)rowb, sc.b.as<uint8_t>(), src->cols, src->rows, src->cols, cn, sc.w.p, s)) return e;
  }
  SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, sc.a.p, rowb, rowb, src->rows,
                             dst->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

int ssk_median_filter_bad_pixels(ssk_mat *image, double variation_threshold) {
  return bp_filter(image, variation_threshold, false, "median_filter_bad_pixels");
}

int ssk_bayer_denoise(ssk_mat *image, double variation_threshold) {
  return bp_filter(image, variation_threshold, true, "bayer_denoise");
}

}  // extern "C"
