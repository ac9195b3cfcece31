// C ABI (include/ssk.h) over the engine classes.  No torch / OpenCV types cross this boundary.
#include <cmath>
#include <cstring>
#include <memory>
#include <new>
#include "ssk_engine.cuh"
#include "ssk_eccflow.cuh"

#include <cfloat>

using namespace ssk;

static constexpr double CV_PI_D = 3.1415926535897932384626433832795;

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
namespace {

int check_mat(const ssk_mat *m, const char *what) {
  if (!m || !m->data || m->rows <= 0 || m->cols <= 0) { set_error(std::string(what) + ": empty image"); return SSK_ERR_INVALID; }
  const int d = type_depth(m->type), cn = type_cn(m->type);
  if (!depth_bytes(d) || cn < 1 || cn > 4) { set_error(std::string(what) + ": unsupported type"); return SSK_ERR_INVALID; }
  if (m->step < (int64_t)m->cols * cn * depth_bytes(d)) { set_error(std::string(what) + ": step smaller than a row"); return SSK_ERR_INVALID; }
  return SSK_OK;
}

inline bool on_device(const ssk_mat *m) { return !m || m->mem == SSK_MEM_DEVICE; }

// entry check of the calls that end with a host wait: a pending stream-ordered chain is waited for first
int ensure_device_ordered();
int ensure_device() {
  if (int e = ensure_device_ordered()) return e;
  return chain_drain();
}
int ensure_device_ordered() {
  static thread_local bool checked = false;
  if (checked) return SSK_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    set_error(std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
    cudaGetLastError();
    return SSK_ERR_CUDA;
  }
  checked = true;
  return SSK_OK;
}

// make `m` available on the device as a dense image; host data goes through `staging`
int to_device(const ssk_mat *m, DevBuf &staging, cudaStream_t s, Img *out, int bpp) {
  const int d = type_depth(m->type), cn = type_cn(m->type);
  out->rows = m->rows; out->cols = m->cols; out->depth = d; out->cn = cn; out->scale = bpp_scale(d, bpp);
  if (m->mem == SSK_MEM_DEVICE) {
    out->data = m->data; out->step = m->step;
    return SSK_OK;
  }
  const size_t rowb = (size_t)m->cols * cn * depth_bytes(d);
  if (int e = staging.ensure(rowb * m->rows)) return e;
  SSK_CUDA(cudaMemcpy2DAsync(staging.p, rowb, m->data, m->step, rowb, m->rows, cudaMemcpyHostToDevice, s));
  out->data = staging.p; out->step = (int64_t)rowb;
  return SSK_OK;
}

// copy a dense device image into a caller buffer (host or device)
int from_device(const void *dsrc, size_t rowb, int rows, ssk_mat *dst, cudaStream_t s) {
  SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, dsrc, rowb, rowb, rows,
                             dst->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
  return SSK_OK;
}

// c_bayer_average::get_acc_counters: cv::multiply(_counter, cv::Scalar(1, 0.5, 1), accw) (c_frame_accumulation.cc:1239-1249)
__global__ void k_bayer_counters(const float *cntr, float *out, int64_t n3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n3) out[i] = (i % 3 == 1) ? cntr[i] * 0.5f : cntr[i];
}

__global__ void k_create_remap(MapCoef m, int rows, int cols, float2 *dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  float u, v;
  map_xy(m, (float)x, (float)y, u, v);
  dst[(int64_t)y * cols + x] = make_float2(u, v);
}

int create_remap_dense(const ssk_transform &t, int rows, int cols, float2 *d, cudaStream_t s) {
  dim3 grid(div_up(cols, 32), div_up(rows, 8));
  k_create_remap<<<grid, 256, 0, s>>>(make_mapcoef(t), rows, cols, d);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

void fill_status(const EccFrame &f, const Ecch &e, ssk_ecc_status *st) {
  if (!st) return;
  st->rho = f.rho; st->min_rho = e.min_rho; st->eps = f.eps;
  st->num_iterations = f.num_iterations; st->max_iterations = e.opts.max_iterations;
  st->ok = f.ok; st->failed = f.failed;
}

// thread-local scratch for the stateless entry points
struct Scratch {
  cudaStream_t stream = nullptr;
  DevBuf a, b, c, d, e;
  int init() {
    if (!stream) SSK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return SSK_OK;
  }
};
Scratch &scratch() { static thread_local Scratch s; return s; }

// the general remap: shared by ssk_remap and ssk_reg_remap
int do_remap(cudaStream_t s, DevBuf &st_src, DevBuf &st_map, DevBuf &st_mask, DevBuf &st_out, DevBuf &st_tmp,
             const ssk_transform *t, const ssk_mat *rmap, const ssk_mat *src, ssk_mat *dst, const ssk_mat *src_mask,
             ssk_mat *dst_mask, int interp, int border, const double bv[4]) {
  interp = remap_interp(interp);
  Tables tab;
  if (int e = get_tables(&tab)) return e;
  SSK_REQUIRE(t || rmap, "remap: either a transform or an explicit map is required");
  int rows, cols;
  Img map_img = {};
  const float2 *d_rmap = nullptr;
  int64_t rmap_step = 0;
  if (rmap) {
    if (int e = check_mat(rmap, "remap map")) return e;
    SSK_REQUIRE(rmap->type == SSK_32FC2, "remap: map must be CV_32FC2");
    if (int e = to_device(rmap, st_map, s, &map_img, 0)) return e;
    d_rmap = static_cast<const float2 *>(map_img.data);
    rmap_step = map_img.step;
    rows = rmap->rows; cols = rmap->cols;
  } else {
    const ssk_mat *g = dst ? dst : dst_mask;
    SSK_REQUIRE(g, "remap: no output requested");
    rows = g->rows; cols = g->cols;
  }
  const MapCoef mc = t ? make_mapcoef(*t) : MapCoef();
  if (dst) {
    if (int e = check_mat(src, "remap src")) return e;
    if (int e = check_mat(dst, "remap dst")) return e;
    SSK_REQUIRE(type_depth(src->type) == SSK_32F && dst->type == src->type, "remap: CV_32F source and destination of the same type");
    SSK_REQUIRE(dst->rows == rows && dst->cols == cols, "remap: dst size must equal the map size");
    Img sim;
    if (int e = to_device(src, st_src, s, &sim, 0)) return e;
    const size_t rowb = (size_t)cols * sim.cn * 4;
    RemapArgs a = {};
    a.src = sim; a.rows = rows; a.cols = cols; a.map = mc; a.rmap = d_rmap; a.rmap_step = rmap_step;
    a.interp = interp; a.border = border;
    for (int i = 0; i < 4; ++i) a.bval[i] = bv ? (float)bv[i] : 0.f;
    if (dst->mem == SSK_MEM_DEVICE) {
      a.dst = static_cast<float *>(dst->data); a.dst_step = dst->step;
      if (int e = launch_remap(a, tab, s)) return e;
    } else {
      if (int e = st_out.ensure(rowb * rows)) return e;
      if (border == SSK_BORDER_TRANSPARENT)   // dst content is an input in this mode
        SSK_CUDA(cudaMemcpy2DAsync(st_out.p, rowb, dst->data, dst->step, rowb, rows, cudaMemcpyHostToDevice, s));
      a.dst = st_out.as<float>(); a.dst_step = (int64_t)rowb;
      if (int e = launch_remap(a, tab, s)) return e;
      if (int e = from_device(st_out.p, rowb, rows, dst, s)) return e;
    }
  }
  if (dst_mask) {
    if (int e = check_mat(dst_mask, "remap dst_mask")) return e;
    SSK_REQUIRE(dst_mask->type == SSK_8UC1 && dst_mask->rows == rows && dst_mask->cols == cols, "remap: dst_mask must be CV_8UC1 of the map size");
    RemapMaskArgs m = {};
    Img mim = {};
    if (src_mask) {
      if (int e = check_mat(src_mask, "remap src_mask")) return e;
      SSK_REQUIRE(src_mask->type == SSK_8UC1, "remap: src_mask must be CV_8UC1");
      if (int e = to_device(src_mask, st_mask, s, &mim, 0)) return e;
      m.src_mask = static_cast<const uint8_t *>(mim.data); m.src_mask_step = mim.step;
      m.src_rows = src_mask->rows; m.src_cols = src_mask->cols;
    } else {
      m.src_rows = src ? src->rows : rows; m.src_cols = src ? src->cols : cols;
    }
    if (int e = st_tmp.ensure((size_t)rows * cols * 2)) return e;
    m.tmp = st_tmp.as<uint8_t>();
    uint8_t *dense = st_tmp.as<uint8_t>() + (size_t)rows * cols;
    m.rows = rows; m.cols = cols; m.map = mc; m.rmap = d_rmap; m.rmap_step = rmap_step; m.interp = interp;
    if (dst_mask->mem == SSK_MEM_DEVICE) { m.dst = static_cast<uint8_t *>(dst_mask->data); m.dst_step = dst_mask->step; }
    else { m.dst = dense; m.dst_step = cols; }
    if (int e = launch_remap_mask(m, tab, s)) return e;
    if (dst_mask->mem != SSK_MEM_DEVICE)
      if (int e = from_device(dense, (size_t)cols, rows, dst_mask, s)) return e;
  }
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// handle types
// ------------------------------------------------------------------------------------------------

extern "C" {

void ssk_ecch_options_default(ssk_ecch_options *o) {
  o->epsx = 1e-5; o->reference_smooth_sigma = 1; o->input_smooth_sigma = 1; o->update_step_scale = 1;
  o->method = SSK_ECC_INVERSE_COMPOSITIONAL_LM; o->interpolation = SSK_INTER_LINEAR;
  o->max_iterations = 50; o->minimum_image_size = 8; o->maxlevel = 0;
}

void ssk_registration_options_default(ssk_registration_options *o) {
  memset(o, 0, sizeof(*o));
  o->motion_type = SSK_MOTION_AFFINE;
  o->interpolation = SSK_INTER_LINEAR;
  o->border_mode = SSK_BORDER_REFLECT101;
  o->ecc.scale = 0.5; o->ecc.eps = 0.2; o->ecc.min_rho = 0.8;
  o->ecc.input_smooth_sigma = 1.0; o->ecc.reference_smooth_sigma = 1.0; o->ecc.update_step_scale = 1.5;
  o->ecc.se_radius = 5; o->ecc.ecc_method = SSK_ECC_LM; o->ecc.max_iterations = 50;
  o->ecc.ecch_max_level = 0; o->ecc.ecch_minimum_image_size = 16;
  o->ecc.normalization_noise = 0.01; o->ecc.normalization_scale = 0;
  o->ecc.ecch_estimate_translation_first = 1; o->ecc.replace_planetary_disk_with_mask = 0;
  o->enable_ecc_registration = 0;   // reference default (c_frame_registration.h:133); callers of this path set it
  o->enable_eccflow_registration = 0;
  ssk_eccflow_registration_options_default(&o->eccflow);   // c_frame_registration.h:88-100
}

int ssk_transform_init(ssk_transform *t, int motion_type) { return make_transform(t, motion_type); }
int ssk_transform_scale(ssk_transform *t, double factor) { return host_scale_transform(t, factor); }

int ssk_transform_create_remap(const ssk_transform *t, int rows, int cols, ssk_mat *rmap) {
  if (int e = ensure_device()) return e;
  if (int e = check_mat(rmap, "create_remap")) return e;
  SSK_REQUIRE(rmap->type == SSK_32FC2 && rmap->rows == rows && rmap->cols == cols, "create_remap: rmap must be CV_32FC2 rows x cols");
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  if (int e = sc.a.ensure((size_t)rows * cols * 8)) return e;
  if (int e = create_remap_dense(*t, rows, cols, sc.a.as<float2>(), sc.stream)) return e;
  if (int e = from_device(sc.a.p, (size_t)cols * 8, rows, rmap, sc.stream)) return e;
  SSK_CUDA(cudaStreamSynchronize(sc.stream));
  return SSK_OK;
}

int ssk_remap(const ssk_transform *t, const ssk_mat *rmap, const ssk_mat *src, ssk_mat *dst, const ssk_mat *src_mask,
              ssk_mat *dst_mask, int interpolation, int border_mode, const double border_value[4]) {
  if (int e = ensure_device()) return e;
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  return do_remap(sc.stream, sc.a, sc.b, sc.c, sc.d, sc.e, t, rmap, src, dst, src_mask, dst_mask, interpolation,
                  border_mode, border_value);
}

// ---- c_ecch ------------------------------------------------------------------------------------
int ssk_ecch_create(const ssk_ecch_options *opts, ssk_ecch **out) {
  if (int e = ensure_device()) return e;
  SSK_REQUIRE(opts && out, "ssk_ecch_create: null argument");
  ssk_ecch *h = new (std::nothrow) ssk_ecch();
  SSK_REQUIRE(h, "out of memory");
  cudaError_t ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete h; return cuda_fail(ce, "cudaStreamCreate", __FILE__, __LINE__); }
  h->e.init(*opts, h->stream);
  *out = h;
  return SSK_OK;
}

int ssk_ecch_destroy(ssk_ecch *h) { delete h; return SSK_OK; }

static int ecch_image_to_gray(ssk_ecch *h, const ssk_mat *image, float *d_dst) {
  // ecc_convert_input_image (ecc2.cc:345-382): convertTo(CV_32F) without scaling; colour -> cvtColor(BGR2GRAY)
  Img im;
  if (int e = to_device(image, h->staging, h->stream, &im, 0)) return e;
  im.scale = 1.f;
  SSK_REQUIRE(im.cn == 1 || im.cn == 3, "c_ecch: 1 or 3 channel images");
  return launch_to_gray(im, nullptr, d_dst, nullptr, 1, h->stream);
}

int ssk_ecch_set_reference_image(ssk_ecch *h, const ssk_mat *image, const ssk_mat *mask) {
  SSK_REQUIRE(h, "null handle");
  if (int e = check_mat(image, "c_ecch::set_reference_image")) return e;
  if (int e = h->d_ptr.ensure((size_t)image->rows * image->cols * 4)) return e;
  if (int e = ecch_image_to_gray(h, image, h->d_ptr.as<float>())) return e;
  const uint8_t *d_mask = nullptr;
  if (mask) {
    int64_t step = 0;
    const uint8_t *dm = nullptr;
    if (int e = mask_to_device(mask, image->rows, image->cols, h->st_mask, h->stream, &dm, &step)) return e;
    if (step != image->cols) {     // Ecch wants a dense mask
      if (int e = h->st_mask2.ensure((size_t)image->rows * image->cols)) return e;
      SSK_CUDA(cudaMemcpy2DAsync(h->st_mask2.p, image->cols, dm, step, image->cols, image->rows, cudaMemcpyDeviceToDevice, h->stream));
      dm = h->st_mask2.as<uint8_t>();
    }
    d_mask = dm;
  }
  if (int e = h->e.set_reference(h->d_ptr.as<float>(), image->rows, image->cols, d_mask)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

int ssk_ecch_align(ssk_ecch *h, const ssk_mat *image, const ssk_mat *mask, ssk_transform *t, ssk_ecc_status *status) {
  SSK_REQUIRE(h && t, "null argument");
  SSK_REQUIRE(h->e.have_reference, "c_ecch: no reference image was set");
  if (int e = check_mat(image, "c_ecch::align")) return e;
  SSK_REQUIRE(image->rows == h->e.lh[0] && image->cols == h->e.lw[0], "c_ecch: current image size differs from the reference image size");
  if (int e = h->e.reserve(1)) return e;
  if (int e = ecch_image_to_gray(h, image, h->e.level0_scratch(0))) return e;
  if (int e = h->e.prepare_current(h->e.level0_scratch_ptrs(), 1)) return e;
  if (mask) {
    const uint8_t *d_mask = nullptr;
    int64_t mstep = 0;
    if (int e = mask_to_device(mask, image->rows, image->cols, h->st_mask2, h->stream, &d_mask, &mstep)) return e;
    if (mstep != image->cols) {     // a strided device mask: the pyramid builder wants it dense
      if (int e = h->st_mask2.ensure((size_t)image->rows * image->cols)) return e;
      SSK_CUDA(cudaMemcpy2DAsync(h->st_mask2.p, image->cols, d_mask, mstep, image->cols, image->rows, cudaMemcpyDeviceToDevice, h->stream));
      d_mask = h->st_mask2.as<uint8_t>();
    }
    if (int e = h->e.prepare_current_mask(d_mask)) return e;
  }
  h->e.translation_first = 0; h->e.check_rho = 0; h->e.final_scale = 1.0;
  if (int e = h->e.align(1, *t)) return e;
  if (int e = h->e.download_frames(1)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  const EccFrame &f = h->e.host_frames()[0];
  *t = f.t;
  fill_status(f, h->e, status);
  return SSK_OK;
}

int ssk_ecch_set_trace(ssk_ecch *h, int max_records) {
  SSK_REQUIRE(h, "null handle");
  return h->e.enable_trace(max_records);
}

int ssk_ecch_get_trace(ssk_ecch *h, float *records, int max_records, int *n) {
  SSK_REQUIRE(h && records && n, "null argument");
  return h->e.fetch_trace(records, max_records, n);
}

int ssk_reg_set_trace(ssk_reg *h, int max_records) {
  SSK_REQUIRE(h, "null handle");
  return h->r.ecch.enable_trace(max_records);
}

int ssk_reg_get_trace(ssk_reg *h, float *records, int max_records, int *n) {
  SSK_REQUIRE(h && records && n, "null argument");
  return h->r.ecch.fetch_trace(records, max_records, n);
}

int ssk_ecch_num_levels(const ssk_ecch *h) { return h ? h->e.nlevels : 0; }

int ssk_ecch_level_size(const ssk_ecch *h, int level, int *cols, int *rows) {
  SSK_REQUIRE(h && level >= 0 && level < h->e.nlevels, "c_ecch: bad level");
  *cols = h->e.lw[level]; *rows = h->e.lh[level];
  return SSK_OK;
}

int ssk_ecch_get_image(const ssk_ecch *h, int which, int level, ssk_mat *dst) {
  SSK_REQUIRE(h && level >= 0 && level < h->e.nlevels, "c_ecch: bad level");
  if (int e = check_mat(dst, "c_ecch::get_image")) return e;
  SSK_REQUIRE(dst->type == SSK_32FC1 && dst->rows == h->e.lh[level] && dst->cols == h->e.lw[level], "c_ecch::get_image: dst must be CV_32FC1 of the level size");
  SSK_REQUIRE(which == 0 || h->e.capacity > 0, "c_ecch: no current image");
  const float *src = which == 0 ? h->e.reference_level(level) : h->e.current_level(0, level);
  if (int e = from_device(src, (size_t)h->e.lw[level] * 4, h->e.lh[level], dst, h->stream)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  return SSK_OK;
}

// ---- c_frame_registration ----------------------------------------------------------------------
int ssk_reg_create(const ssk_registration_options *opts, ssk_reg **out) {
  if (int e = ensure_device()) return e;
  SSK_REQUIRE(opts && out, "ssk_reg_create: null argument");
  ssk_reg *h = new (std::nothrow) ssk_reg();
  SSK_REQUIRE(h, "out of memory");
  cudaStream_t s;
  cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete h; return cuda_fail(ce, "cudaStreamCreate", __FILE__, __LINE__); }
  if (int e = h->r.init(*opts, s, true)) { delete h; return e; }
  *out = h;
  return SSK_OK;
}

int ssk_reg_destroy(ssk_reg *h) { delete h; return SSK_OK; }

int ssk_reg_setup_reference_frame(ssk_reg *h, const ssk_mat *image, const ssk_mat *mask, int bpp) {
  SSK_REQUIRE(h, "null handle");
  if (int e = check_mat(image, "setup_reference_frame")) return e;
  Img im;
  if (int e = to_device(image, h->staging, h->r.stream, &im, bpp)) return e;
  SSK_REQUIRE(im.cn == 1 || im.cn == 3, "c_frame_registration: 1 or 3 channel frames");
  const uint8_t *d_mask = nullptr;
  int64_t mstep = 0;
  if (mask) { if (int e = mask_to_device(mask, image->rows, image->cols, h->st_mask, h->r.stream, &d_mask, &mstep)) return e; }
  if (int e = h->r.setup_reference(im, d_mask, mstep)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->r.stream));
  return SSK_OK;
}

int ssk_reg_register_frame(ssk_reg *h, const ssk_mat *image, const ssk_mat *mask, int bpp, ssk_transform *t_out,
                           ssk_ecc_status *status) {
  SSK_REQUIRE(h, "null handle");
  if (int e = check_mat(image, "register_frame")) return e;
  Img im;
  if (int e = to_device(image, h->staging, h->r.stream, &im, bpp)) return e;
  if (int e = h->d_ptr.ensure(sizeof(void *))) return e;
  const void *p = im.data;
  SSK_CUDA(cudaMemcpyAsync(h->d_ptr.p, &p, sizeof(p), cudaMemcpyHostToDevice, h->r.stream));
  SSK_CUDA(cudaStreamSynchronize(h->r.stream));   // &p is a stack variable
  const uint8_t *d_mask = nullptr;
  int64_t mstep = 0;
  if (mask) { if (int e = mask_to_device(mask, image->rows, image->cols, h->st_mask, h->r.stream, &d_mask, &mstep)) return e; }
  if (int e = h->r.prepare(im, h->d_ptr.as<const void *>(), 1, d_mask, mstep)) return e;
  if (int e = h->r.register_batch(1)) return e;
  if (int e = h->r.ecch.download_frames(1)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->r.stream));
  const EccFrame &f = h->r.ecch.host_frames()[0];
  h->r.current = f.t;
  h->r.have_current = f.ok != 0;
  if (t_out) *t_out = f.t;
  fill_status(f, h->r.ecch, status);
  if (!f.ok) {
    char buf[160];
    snprintf(buf, sizeof(buf), "Poor image correlation after align : rho = %g / %g", f.rho, h->r.opts.ecc.min_rho);
    set_error(buf);
    return SSK_ERR_NOT_REGISTERED;
  }
  return SSK_OK;
}

// _current_remap after the c_eccflow stage (c_frame_registration.cc:900-917) as a dense device CV_32FC2 image
static int reg_flow_remap(ssk_reg *h, ssk_mat *m) {
  const int rows = h->r.ref_rows, cols = h->r.ref_cols;
  if (int e = h->st_flowmap.ensure((size_t)rows * cols * 8)) return e;
  if (int e = h->r.flowh->write_remap(0, h->st_flowmap.as<float2>())) return e;
  m->data = h->st_flowmap.p; m->step = (int64_t)cols * 8; m->rows = rows; m->cols = cols; m->type = SSK_32FC2; m->mem = SSK_MEM_DEVICE;
  return SSK_OK;
}

int ssk_reg_get_current_remap(ssk_reg *h, ssk_mat *rmap) {
  SSK_REQUIRE(h && h->r.have_current, "c_frame_registration: no registered frame");
  if (h->r.flow_enabled()) {
    if (int e = check_mat(rmap, "current_remap")) return e;
    SSK_REQUIRE(rmap->type == SSK_32FC2 && rmap->rows == h->r.ref_rows && rmap->cols == h->r.ref_cols, "current_remap: rmap must be CV_32FC2 of the reference size");
    ssk_mat m;
    if (int e = reg_flow_remap(h, &m)) return e;
    if (int e = from_device(m.data, (size_t)m.cols * 8, m.rows, rmap, h->r.stream)) return e;
    SSK_CUDA(cudaStreamSynchronize(h->r.stream));
    return SSK_OK;
  }
  return ssk_transform_create_remap(&h->r.current, h->r.ref_rows, h->r.ref_cols, rmap);
}

int ssk_reg_remap(ssk_reg *h, const ssk_mat *rmap, const ssk_mat *src, ssk_mat *dst, const ssk_mat *src_mask,
                  ssk_mat *dst_mask, int interpolation, int border_mode, const double border_value[4]) {
  SSK_REQUIRE(h, "null handle");
  SSK_REQUIRE(rmap || h->r.have_current, "c_frame_registration::remap: no current transform");
  // c_frame_registration.cc:1299-1309
  if (interpolation < 0) interpolation = h->r.opts.interpolation;
  const double *bv = border_value;
  if (border_mode < 0) { border_mode = h->r.opts.border_mode; bv = h->r.opts.border_value; }
  if (!rmap && h->r.flow_enabled()) {
    ssk_mat m;
    if (int e = reg_flow_remap(h, &m)) return e;
    return do_remap(h->r.stream, h->staging, h->st_map, h->st_mask, h->st_out, h->st_tmp, nullptr, &m, src, dst, src_mask, dst_mask,
                    interpolation, border_mode, bv);
  }
  return do_remap(h->r.stream, h->staging, h->st_map, h->st_mask, h->st_out, h->st_tmp, rmap ? nullptr : &h->r.current, rmap,
                  src, dst, src_mask, dst_mask, interpolation, border_mode, bv);
}

// ---- c_frame_accumulation ----------------------------------------------------------------------
int ssk_acc_create(int kind, ssk_acc **out) {
  if (int e = ensure_device()) return e;
  SSK_REQUIRE(out && (kind == SSK_ACC_WEIGHTED_AVERAGE || kind == SSK_ACC_BAYER_AVERAGE), "ssk_acc_create: bad kind");
  ssk_acc *h = new (std::nothrow) ssk_acc();
  SSK_REQUIRE(h, "out of memory");
  h->a.kind = kind;
  cudaError_t ce = cudaStreamCreateWithFlags(&h->a.stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete h; return cuda_fail(ce, "cudaStreamCreate", __FILE__, __LINE__); }
  h->a.own_stream = true;
  *out = h;
  return SSK_OK;
}

int ssk_acc_destroy(ssk_acc *h) { delete h; return SSK_OK; }
int ssk_acc_clear(ssk_acc *h) { SSK_REQUIRE(h, "null handle"); if (int e = chain_drain()) return e; return h->a.clear(); }

int ssk_acc_add(ssk_acc *h, const ssk_mat *src, const ssk_mat *weights, int bpp) {
  SSK_REQUIRE(h, "null handle");
  Acc &a = h->a;
  if (int e = check_mat(src, "c_frame_accumulation::add")) return e;
  if (int e = chain_wait(a.stream)) return e;
  Img im;
  if (int e = to_device(src, a.staging, a.stream, &im, bpp)) return e;
  int wtype = -1;
  Img wim = {};
  if (weights) {
    if (int e = check_mat(weights, "c_frame_accumulation::add weights")) return e;
    SSK_REQUIRE(weights->rows == src->rows && weights->cols == src->cols, "frame accumulation: image and weights sizes not match");
    SSK_REQUIRE(weights->type == SSK_8UC1 || weights->type == SSK_32FC1, "frame accumulation: weights must be CV_8UC1 or CV_32FC1");
    wtype = weights->type;
    if (int e = to_device(weights, a.wstaging, a.stream, &wim, 0)) return e;
  }
  if (a.kind == SSK_ACC_BAYER_AVERAGE) {
    SSK_REQUIRE(im.cn == 1, "c_bayer_average: single channel raw frames");
    if (int e = a.ensure(im.rows, im.cols, 3)) return e;
    BayerAccArgs b = {};
    b.src = im; b.have_map = a.have_map ? 1 : 0;
    if (a.have_map) {
      if (a.rmap_explicit) {
        SSK_REQUIRE(a.rmap_rows == im.rows && a.rmap_cols == im.cols, "c_bayer_average: remap size differs from the frame size");
        b.rmap = a.rmap.as<float2>(); b.rmap_step = (int64_t)im.cols * 8;
      }
      else b.map = make_mapcoef(a.map_t);
    }
    b.weights = wim.data; b.w_step = wim.step; b.wtype = wtype; b.colorid = a.colorid;
    b.acc = a.acc.as<float>(); b.cntr = a.wacc.as<float>();
    if (int e = launch_bayer_add(b, a.stream)) return e;
  } else {
    if (int e = a.ensure(im.rows, im.cols, im.cn)) return e;
    AccAddArgs x = {};
    x.src = im; x.weights = wim.data; x.w_step = wim.step; x.wtype = wtype;
    x.acc = a.acc.as<float>(); x.wacc = a.wacc.as<float>();
    if (int e = launch_acc_add(x, a.stream)) return e;
  }
  ++a.frames;
  return chain_finish(a.stream, on_device(src) && on_device(weights));   // host buffers of the caller are free again on return
}

int ssk_acc_compute(ssk_acc *h, ssk_mat *avg, ssk_mat *mask, double dscale) {
  SSK_REQUIRE(h, "null handle");
  if (int e = chain_drain()) return e;
  Acc &a = h->a;
  if (a.frames < 1) { set_error("c_frame_accumulation::compute: no accumulated frames"); return SSK_ERR_STATE; }
  const int ocn = a.kind == SSK_ACC_BAYER_AVERAGE ? 3 : a.cn;
  const size_t rowb = (size_t)a.cols * ocn * 4;
  if (avg) {
    if (int e = check_mat(avg, "compute avg")) return e;
    SSK_REQUIRE(type_depth(avg->type) == SSK_32F && type_cn(avg->type) == ocn && avg->rows == a.rows && avg->cols == a.cols,
                "compute: avg must be CV_32F with the accumulator's size and channels");
  }
  if (mask) {
    if (int e = check_mat(mask, "compute mask")) return e;
    SSK_REQUIRE(mask->type == SSK_8UC1 && mask->rows == a.rows && mask->cols == a.cols, "compute: mask must be CV_8UC1 of the accumulator size");
  }
  if (int e = a.out_staging.ensure(rowb * a.rows + (size_t)a.rows * a.cols)) return e;
  float *d_avg = a.out_staging.as<float>();
  uint8_t *d_mask = reinterpret_cast<uint8_t *>(a.out_staging.as<char>() + rowb * a.rows);
  if (a.kind == SSK_ACC_BAYER_AVERAGE) {
    if (int e = launch_bayer_compute(a.acc.as<float>(), a.wacc.as<float>(), a.rows, a.cols, avg ? d_avg : nullptr, (int64_t)rowb,
                                     mask ? d_mask : nullptr, a.cols, a.stream)) return e;
  } else {
    if (int e = launch_acc_compute(a.acc.as<float>(), a.wacc.as<float>(), a.rows, a.cols, a.cn, (float)dscale, avg ? d_avg : nullptr,
                                   (int64_t)rowb, mask ? d_mask : nullptr, a.cols, a.stream)) return e;
  }
  if (avg) if (int e = from_device(d_avg, rowb, a.rows, avg, a.stream)) return e;
  if (mask) if (int e = from_device(d_mask, (size_t)a.cols, a.rows, mask, a.stream)) return e;
  SSK_CUDA(cudaStreamSynchronize(a.stream));
  return SSK_OK;
}

// c_image_stacking_pipeline.cc:742-767: compute() then average_pyramid_inpaint(avg, mask, avg, mask, max_levels)
int ssk_acc_compute_inpainted(ssk_acc *h, ssk_mat *avg, ssk_mat *mask, double dscale, int max_levels) {
  SSK_REQUIRE(h && avg, "null argument");
  Acc &a = h->a;
  if (a.frames < 1) { set_error("c_frame_accumulation::compute: no accumulated frames"); return SSK_ERR_STATE; }
  const int ocn = a.kind == SSK_ACC_BAYER_AVERAGE ? 3 : a.cn;
  const size_t rowb = (size_t)a.cols * ocn * 4, npx = (size_t)a.rows * a.cols;
  if (int e = check_mat(avg, "compute avg")) return e;
  SSK_REQUIRE(type_depth(avg->type) == SSK_32F && type_cn(avg->type) == ocn && avg->rows == a.rows && avg->cols == a.cols,
              "compute: avg must be CV_32F with the accumulator's size and channels");
  if (mask) {
    if (int e = check_mat(mask, "compute mask")) return e;
    SSK_REQUIRE(mask->type == SSK_8UC1 && mask->rows == a.rows && mask->cols == a.cols, "compute: mask must be CV_8UC1 of the accumulator size");
  }
  const size_t img_b = (rowb * a.rows + 15) & ~(size_t)15, msk_b = (npx + 15) & ~(size_t)15;
  if (int e = a.out_staging.ensure(2 * (img_b + msk_b) + inpaint_work_bytes(a.rows, a.cols, ocn, max_levels))) return e;
  char *base = a.out_staging.as<char>();
  float *d_avg = reinterpret_cast<float *>(base), *d_out = reinterpret_cast<float *>(base + img_b);
  uint8_t *d_mask = reinterpret_cast<uint8_t *>(base + 2 * img_b), *d_omask = d_mask + msk_b;
  void *work = base + 2 * (img_b + msk_b);
  if (a.kind == SSK_ACC_BAYER_AVERAGE) {
    if (int e = launch_bayer_compute(a.acc.as<float>(), a.wacc.as<float>(), a.rows, a.cols, d_avg, (int64_t)rowb, d_mask, a.cols, a.stream)) return e;
  } else {
    if (int e = launch_acc_compute(a.acc.as<float>(), a.wacc.as<float>(), a.rows, a.cols, a.cn, (float)dscale, d_avg, (int64_t)rowb,
                                   d_mask, a.cols, a.stream)) return e;
  }
  int full = 0;
  if (int e = launch_average_pyramid_inpaint(d_avg, (int64_t)rowb, d_mask, a.cols, a.rows, a.cols, ocn, max_levels, work, d_out, d_omask,
                                             &full, a.stream)) return e;
  if (int e = from_device(d_out, rowb, a.rows, avg, a.stream)) return e;
  if (mask) if (int e = from_device(d_omask, (size_t)a.cols, a.rows, mask, a.stream)) return e;
  SSK_CUDA(cudaStreamSynchronize(a.stream));
  return SSK_OK;
}

int ssk_acc_get_counters(ssk_acc *h, ssk_mat *accw) {
  SSK_REQUIRE(h, "null handle");
  if (int e = chain_drain()) return e;
  Acc &a = h->a;
  SSK_REQUIRE(a.wacc.p, "get_acc_counters: empty accumulator");
  if (int e = check_mat(accw, "get_acc_counters")) return e;
  const int wcn = a.kind == SSK_ACC_BAYER_AVERAGE ? 3 : 1;
  SSK_REQUIRE(type_depth(accw->type) == SSK_32F && type_cn(accw->type) == wcn && accw->rows == a.rows && accw->cols == a.cols,
              "get_acc_counters: CV_32F buffer of the accumulator size (3 channels for bayer)");
  const void *src = a.wacc.p;
  if (a.kind == SSK_ACC_BAYER_AVERAGE) {
    const int64_t n3 = (int64_t)a.rows * a.cols * 3;
    if (int e = a.out_staging.ensure((size_t)n3 * 4)) return e;
    k_bayer_counters<<<(unsigned)((n3 + 255) / 256), 256, 0, a.stream>>>(a.wacc.as<float>(), a.out_staging.as<float>(), n3);
    SSK_LAUNCH_CHECK();
    src = a.out_staging.p;
  }
  if (int e = from_device(src, (size_t)a.cols * wcn * 4, a.rows, accw, a.stream)) return e;
  SSK_CUDA(cudaStreamSynchronize(a.stream));
  return SSK_OK;
}

int ssk_acc_reinitialize(ssk_acc *h, const ssk_mat *src, const ssk_mat *accw) {
  SSK_REQUIRE(h, "null handle");
  Acc &a = h->a;
  SSK_REQUIRE(a.kind == SSK_ACC_WEIGHTED_AVERAGE, "c_bayer_average::reinitialize returns false in the reference");
  if (int e = check_mat(src, "reinitialize src")) return e;
  if (int e = check_mat(accw, "reinitialize accw")) return e;
  SSK_REQUIRE(type_depth(src->type) == SSK_32F && accw->type == SSK_32FC1 && accw->rows == src->rows && accw->cols == src->cols,
              "reinitialize: CV_32F image and CV_32FC1 weights of the same size");
  a.clear();
  if (int e = a.ensure(src->rows, src->cols, type_cn(src->type))) return e;
  const size_t rowb = (size_t)src->cols * a.cn * 4;
  const cudaMemcpyKind k1 = src->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind k2 = accw->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  SSK_CUDA(cudaMemcpy2DAsync(a.acc.p, rowb, src->data, src->step, rowb, src->rows, k1, a.stream));
  SSK_CUDA(cudaMemcpy2DAsync(a.wacc.p, (size_t)src->cols * 4, accw->data, accw->step, (size_t)src->cols * 4, src->rows, k2, a.stream));
  SSK_CUDA(cudaStreamSynchronize(a.stream));
  a.frames = 1;
  return SSK_OK;
}

int ssk_acc_size(const ssk_acc *h, int *cols, int *rows, int *channels) {
  SSK_REQUIRE(h, "null handle");
  if (cols) *cols = h->a.cols;
  if (rows) *rows = h->a.rows;
  if (channels) *channels = h->a.kind == SSK_ACC_BAYER_AVERAGE && h->a.rows ? 3 : h->a.cn;
  return SSK_OK;
}

int ssk_acc_frames(const ssk_acc *h) { return h ? h->a.frames : 0; }

int ssk_acc_set_bayer_pattern(ssk_acc *h, int colorid) {
  SSK_REQUIRE(h, "null handle");
  SSK_REQUIRE(colorid >= SSK_COLORID_BAYER_RGGB && colorid <= SSK_COLORID_BAYER_BGGR, "set_bayer_pattern: RGGB/GRBG/GBRG/BGGR");
  h->a.colorid = colorid;
  return SSK_OK;
}

int ssk_acc_set_remap(ssk_acc *h, const ssk_transform *t, const ssk_mat *rmap) {
  SSK_REQUIRE(h, "null handle");
  Acc &a = h->a;
  if (!t && !rmap) { a.have_map = false; a.rmap_explicit = false; return SSK_OK; }
  if (rmap) {
    if (int e = check_mat(rmap, "set_remap")) return e;
    SSK_REQUIRE(rmap->type == SSK_32FC2, "set_remap: CV_32FC2 map");
    const size_t rowb = (size_t)rmap->cols * 8;
    if (int e = a.rmap.ensure(rowb * rmap->rows)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(a.rmap.p, rowb, rmap->data, rmap->step, rowb, rmap->rows,
                               rmap->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, a.stream));
    SSK_CUDA(cudaStreamSynchronize(a.stream));
    a.rmap_explicit = true;
    a.rmap_rows = rmap->rows; a.rmap_cols = rmap->cols;
  } else {
    a.map_t = *t;
    a.rmap_explicit = false;
  }
  a.have_map = true;
  return SSK_OK;
}

int ssk_acc_device_state(ssk_acc *h, void **acc, void **weights, int64_t *acc_bytes, int64_t *weights_bytes) {
  SSK_REQUIRE(h && h->a.acc.p, "device_state: empty accumulator");
  Acc &a = h->a;
  const int64_t npix = (int64_t)a.rows * a.cols;
  if (acc) *acc = a.acc.p;
  if (weights) *weights = a.wacc.p;
  if (acc_bytes) *acc_bytes = npix * (a.kind == SSK_ACC_BAYER_AVERAGE ? 3 : a.cn) * 4;
  if (weights_bytes) *weights_bytes = npix * (a.kind == SSK_ACC_BAYER_AVERAGE ? 3 : 1) * 4;
  return SSK_OK;
}

int ssk_acc_to_sum_form(ssk_acc *h) {
  SSK_REQUIRE(h && h->a.acc.p, "to_sum_form: empty accumulator");
  Acc &a = h->a;
  if (a.kind == SSK_ACC_BAYER_AVERAGE) return SSK_OK;   // already sums
  if (int e = launch_acc_sum_form(a.acc.as<float>(), a.wacc.as<float>(), (int64_t)a.rows * a.cols, a.cn, 1, a.stream)) return e;
  SSK_CUDA(cudaStreamSynchronize(a.stream));
  return SSK_OK;
}

int ssk_acc_from_sum_form(ssk_acc *h, int accumulated_frames) {
  SSK_REQUIRE(h && h->a.acc.p, "from_sum_form: empty accumulator");
  Acc &a = h->a;
  if (a.kind != SSK_ACC_BAYER_AVERAGE) {
    if (int e = launch_acc_sum_form(a.acc.as<float>(), a.wacc.as<float>(), (int64_t)a.rows * a.cols, a.cn, 0, a.stream)) return e;
    SSK_CUDA(cudaStreamSynchronize(a.stream));
  }
  a.frames = accumulated_frames;
  return SSK_OK;
}

// ---- weight maps ---------------------------------------------------------------------------------
int ssk_local_variance_map(const ssk_mat *image, int bpp, int dscale, int kradius, int uscale, ssk_mat *map, double *Q) {
  if (int e = ensure_device()) return e;
  if (int e = check_mat(image, "compute_local_variance_map")) return e;
  SSK_REQUIRE(uscale >= 0 && uscale <= 12, "compute_local_variance_map: uscale 0..12");
  SSK_REQUIRE(dscale >= 0 && dscale <= 6, "compute_local_variance_map: dscale 0..6");
  if (map) {
    if (int e = check_mat(map, "compute_local_variance_map map")) return e;
    SSK_REQUIRE(map->type == SSK_32FC1 && map->rows == image->rows && map->cols == image->cols, "map must be CV_32FC1 of the image size");
  }
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  Img im;
  if (int e = to_device(image, sc.a, s, &im, bpp < 0 ? 0 : bpp)) return e;
  // bpp < 0: the image is taken as the reference's own overload takes it (select_master_frame ranks raw integer frames,
  // c_image_stacking_pipeline_base.cc:376): depth_scale = 20 / maxval(depth) (c_local_variance_sharpness_measure.cc:203-209),
  // i.e. integer samples normalised by 1 / 255 or 1 / 65535.  The pdownscale chain runs in float here, where cv::pyrDown
  // rounds an integer image at every level: the metric agrees to ~1e-4 relative on integer frames, exactly on CV_32F.
  if (bpp < 0) im.scale = im.depth == SSK_8U ? (float)(1.0 / 255.0) : im.depth == SSK_16U ? (float)(1.0 / 65535.0) : 1.f;
  SSK_REQUIRE(im.cn == 1 || im.cn == 3, "compute_local_variance_map: 1 or 3 channels");
  // pdownscale (c_local_variance_sharpness_measure.cc:28-52)
  const size_t n = (size_t)im.rows * im.cols;
  if (int e = sc.b.ensure(n * 4 * 2)) return e;
  float *bufA = sc.b.as<float>(), *bufB = bufA + n;
  int r = im.rows, c = im.cols;
  const float *M = nullptr;
  if (dscale > 0 && std::min(r, c) >= 4) {
    Img cur = im;
    float *dst = bufA;
    for (int l = 0; l < dscale; ++l) {
      const int nr = (cur.rows + 1) / 2, nc = (cur.cols + 1) / 2;
      PyrDownArgs pd = {};
      pd.src = cur; pd.dst = dst; pd.dst_rows = nr; pd.dst_cols = nc; pd.batch = 1; pd.post_scale = 1.f;
      if (int e = launch_pyrdown(pd, s)) return e;
      cur.data = dst; cur.step = (int64_t)nc * 4; cur.rows = nr; cur.cols = nc; cur.depth = SSK_32F; cur.cn = 1; cur.scale = 1.f;
      M = dst;
      dst = dst == bufA ? bufB : bufA;
      if (std::min(nr, nc) < 4) break;
    }
    r = cur.rows; c = cur.cols;
  } else {
    if (int e = launch_to_gray(im, nullptr, bufA, nullptr, 1, s)) return e;
    M = bufA;
  }
  const int nb = w1_num_blocks(r, c);
  if (int e = sc.c.ensure((size_t)r * c * 4 * 2)) return e;      // gmap + uscale scratch
  if (int e = sc.d.ensure((size_t)nb * 2 * 8 + 4 * 8 + (size_t)(im.rows + im.cols + 8) * sizeof(int2) + 16)) return e;
  if (int e = sc.e.ensure(n * 4)) return e;
  W1Args w = {};
  w.M = M; w.rows = r; w.cols = c; w.kradius = std::max(1, kradius); w.depth_scale = 20.0;
  w.gmap = sc.c.as<float>(); w.partials = sc.d.as<double>(); w.stats = sc.d.as<double>() + (size_t)nb * 2;
  w.out = map ? sc.e.as<float>() : nullptr; w.full_rows = im.rows; w.full_cols = im.cols; w.batch = 1;
  w.axis_tab = reinterpret_cast<int2 *>(sc.d.as<double>() + (size_t)nb * 2 + 4);
  w.uscale = uscale; w.gmap2 = sc.c.as<float>() + (size_t)r * c;
  if (int e = launch_w1(w, s)) return e;
  double stats[4];
  SSK_CUDA(cudaMemcpyAsync(stats, w.stats, sizeof(stats), cudaMemcpyDeviceToHost, s));
  if (map) if (int e = from_device(sc.e.p, (size_t)im.cols * 4, im.rows, map, s)) return e;
  SSK_CUDA(cudaStreamSynchronize(s));
  if (Q) *Q = stats[2];
  if (!(stats[0] > 0)) { set_error("compute_local_variance_map: flat image (sum of gradients is zero), map released"); return SSK_ERR_STATE; }
  return SSK_OK;
}

}  // extern "C"

// lpg on a device image; the result (dense, image size) lives in sc.b until the next call that uses sc.b
// reduce_color_channels(s, s, cv::REDUCE_AVG) (lpg.cc:246-248, reduce_channels.cc:11-29): cv::reduce along the channel
// axis of the CV_32F image = float sum in channel order, times (float)(1 / cn)  (checked against cv2.reduce)
template <int DEPTH>
__global__ void __launch_bounds__(256) k_channel_avg(const Img im, float *dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= im.cols || y >= im.rows) return;
  float sum = load_px<DEPTH>(im, y, x, 0);
  for (int c = 1; c < im.cn; ++c) sum = __fadd_rn(sum, load_px<DEPTH>(im, y, x, c));
  dst[(int64_t)y * im.cols + x] = __fmul_rn(sum, (float)(1.0 / im.cn));
}

// pdownscale (lpg.cc:132-156): `level` pyrDowns between the two scratch buffers, stopping once a side drops below 4
static int lpg_pdown(cudaStream_t s, float *bufA, float *bufB, Img cur, int level, float post, float **out, int *orows, int *ocols) {
  float *dst = static_cast<const void *>(cur.data) == bufA ? bufB : bufA;
  if (std::min(cur.rows, cur.cols) < 4) { *out = nullptr; return SSK_OK; }
  for (int l = 0; l < level; ++l) {
    const int nr = (cur.rows + 1) / 2, nc = (cur.cols + 1) / 2;
    PyrDownArgs pd = {};
    pd.src = cur; pd.dst = dst; pd.dst_rows = nr; pd.dst_cols = nc; pd.batch = 1; pd.post_scale = 1.f;
    if (int e = launch_pyrdown(pd, s)) return e;
    cur.data = dst; cur.step = (int64_t)nc * 4; cur.rows = nr; cur.cols = nc; cur.depth = SSK_32F; cur.cn = 1; cur.scale = 1.f;
    *out = dst;
    dst = dst == bufA ? bufB : bufA;
    if (std::min(nr, nc) < 4) break;
  }
  *orows = cur.rows; *ocols = cur.cols;
  if (post != 1.f) return launch_scale_ipow(*out, (int64_t)cur.rows * cur.cols, post, true, 1, s);
  return SSK_OK;
}

// lpg.cc:262-290 after the 5 x 5 operator: pdownscale by uscale - dscale (times that factor), cv::pow(p), pupscale back
static int lpg_finish(cudaStream_t s, float *M, float *bufA, float *bufB, int r, int c, int im_rows, int im_cols, double p, int dscale,
                      int uscale, bool pow_done, float **out) {
  const int ip = (int)p;
  bool tail_done = false;
  if (uscale > 0 && uscale > dscale) {
    // the pyrDown levels lpg_pdown would run from (r, c): `uscale - dscale` levels, or fewer once a level is smaller than 4
    int nd = 0;
    if (std::min(r, c) >= 4) {
      int rr = r, cc = c;
      while (nd < uscale - dscale) {
        rr = (rr + 1) / 2; cc = (cc + 1) / 2; ++nd;
        if (std::min(rr, cc) < 4) break;
      }
    }
    // levels above `tail_px` pixels are launches of their own; the rest (down, scale, power, up again) is one cluster launch
    // (a larger cut loses: eight CTAs walk a 128-pixel level slower than 148 SMs run its own launch)
    static const int tail_px = getenv("SSK_LPG_TAIL_PX") ? atoi(getenv("SSK_LPG_TAIL_PX")) : 64;   // measured on config #4 (frames/s): no tail 4 900, 64: 5 115, 128: 4 630, 256: 4 480
    int lead = 0;
    {
      int rr = r, cc = c;
      while (lead < nd && std::max(rr, cc) > tail_px) { rr = (rr + 1) / 2; cc = (cc + 1) / 2; ++lead; }
    }
    const bool use_tail = nd - lead >= 1 && nd - lead <= 7 && !pow_done && !getenv("SSK_LPG_NO_TAIL");
    Img cur = {};
    cur.data = M; cur.step = (int64_t)c * 4; cur.rows = r; cur.cols = c; cur.depth = SSK_32F; cur.cn = 1; cur.scale = 1.f;
    if (use_tail) {
      if (lead > 0) {
        float *D = nullptr;
        if (int e = lpg_pdown(s, bufA, bufB, cur, lead, 1.f, &D, &r, &c)) return e;
        M = D;
      }
      float *Q = M == bufA ? bufB : bufA;
      if (int e = launch_lpg_tail(M, Q, r, c, nd - lead, (float)(uscale - dscale), ip >= 2 ? ip : 1, s)) return e;
      tail_done = true;
    } else {
      float *D = nullptr;
      if (int e = lpg_pdown(s, bufA, bufB, cur, uscale - dscale, (float)(uscale - dscale), &D, &r, &c)) return e;
      if (D) M = D;
      else { if (int e = launch_scale_ipow(M, (int64_t)r * c, (float)(uscale - dscale), true, 1, s)) return e; }
    }
  }
  if (!tail_done && !pow_done && ip != 0 && ip != 1) { if (int e = launch_scale_ipow(M, (int64_t)r * c, 1.f, false, ip, s)) return e; }
  // pupscale (lpg.cc:158-182): the (w+1)/2 size chain from the image size down to the map size, walked back by cv::pyrUp
  if (r != im_rows || c != im_cols) {
    int cw[32], chh[32], nl = 0;
    cw[0] = im_cols; chh[0] = im_rows;
    while (true) {
      const int nw = (cw[nl] + 1) / 2, nh = (chh[nl] + 1) / 2;
      if (nw == c && nh == r) break;
      SSK_REQUIRE(nw >= c && nh >= r && nl < 30, "lpg: invalid up-scaling size chain");
      ++nl; cw[nl] = nw; chh[nl] = nh;
    }
    for (int l = nl; l >= 0; --l) {
      float *dst = M == bufA ? bufB : bufA;
      PyrUpArgs pu = {};
      pu.src = M; pu.rows = r; pu.cols = c; pu.dst = dst; pu.dst_rows = chh[l]; pu.dst_cols = cw[l]; pu.batch = 1;
      if (int e = launch_pyrup(pu, s)) return e;
      M = dst; r = chh[l]; c = cw[l];
    }
  }
  *out = M;
  return SSK_OK;
}

// `direct`: a dense device buffer of the image size the caller wants the map in (or null); used when the map comes out of a
// single pass, *out then equals it and no copy follows
static int lpg_device(Scratch &sc, Img im, double k, double p, int dscale, int uscale, float **out, float *direct = nullptr) {
  cudaStream_t s = sc.stream;
  SSK_REQUIRE(im.cn >= 1 && im.cn <= 4, "lpg: 1 to 4 channels");
  // lpg.cc:184-200: integer samples are scaled by 1 / max value of the depth
  im.scale = im.depth == SSK_8U ? (float)(1.0 / 255.0) : im.depth == SSK_16U ? (float)(1.0 / 65535.0) : 1.f;
  const size_t n = (size_t)im.rows * im.cols;
  const float alpha = (float)((float)(k / (k + 1)) * (25.f * 25.f));
  const float beta = (float)((float)(1.0 / (k + 1)) * ((float)(100.0 / 36.0) * (float)(100.0 / 36.0)));
  if (int e = sc.b.ensure(n * 4 * 2)) return e;
  float *bufA = sc.b.as<float>(), *bufB = bufA + n;
  if (dscale <= 0 && std::min(im.rows, im.cols) >= 5) {
    // nothing is scaled before the 5 x 5 operator: channel average, operator and (when no scaling follows either) the power in
    // one pass over the frame
    const bool pow_fused = !(uscale > 0 && uscale > dscale);
    if (pow_fused && direct) {
      if (int e = launch_lpg_fused(im, direct, alpha, beta, 1e-9f, (int)p, s)) return e;
      *out = direct;
      return SSK_OK;
    }
    if (int e = launch_lpg_fused(im, bufA, alpha, beta, 1e-9f, pow_fused ? (int)p : 1, s)) return e;
    return lpg_finish(s, bufA, bufA, bufB, im.rows, im.cols, im.rows, im.cols, p, dscale, uscale, pow_fused, out);
  }
  if (im.cn > 1) {
    if (int e = sc.c.ensure(n * 4)) return e;
    const dim3 grid(div_up(im.cols, 32), div_up(im.rows, 8));
    if (im.depth == SSK_32F) k_channel_avg<SSK_32F><<<grid, 256, 0, s>>>(im, sc.c.as<float>());
    else if (im.depth == SSK_16U) k_channel_avg<SSK_16U><<<grid, 256, 0, s>>>(im, sc.c.as<float>());
    else k_channel_avg<SSK_8U><<<grid, 256, 0, s>>>(im, sc.c.as<float>());
    SSK_LAUNCH_CHECK();
    im.data = sc.c.p; im.step = (int64_t)im.cols * 4; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
  }
  float *S = nullptr;
  int r = im.rows, c = im.cols;
  if (dscale > 0) {
    if (int e = lpg_pdown(s, bufA, bufB, im, dscale, (float)(1.0 / (1 + dscale)), &S, &r, &c)) return e;
  }
  if (!S) {   // no down-scaling: plain float copy of the image
    if (int e = launch_to_gray(im, nullptr, bufA, nullptr, 1, s)) return e;
    S = bufA; r = im.rows; c = im.cols;
    if (dscale > 0) { if (int e = launch_scale_ipow(S, (int64_t)r * c, (float)(1.0 / (1 + dscale)), true, 1, s)) return e; }
  }
  float *M = S == bufA ? bufB : bufA;
  if (int e = launch_lpg5x5(S, r, c, M, alpha, beta, 1e-9f, s)) return e;
  return lpg_finish(s, M, bufA, bufB, r, c, im.rows, im.cols, p, dscale, uscale, false, out);
}

extern "C" {

// lpg (core/proc/lpg.cc:223-290): float conversion by 1/maxval(depth), channel average, pdownscale(dscale) * 1/(1+dscale),
// compute_lpg_5x5(k/(k+1), 1/(k+1), 1e-9), pdownscale(uscale - dscale) * (uscale - dscale), pow(p), pyrUp chain back
// to the image size.
int ssk_lpg(const ssk_mat *image, double k, double p, int dscale, int uscale, ssk_mat *map) {
  if (int e = ensure_device_ordered()) return e;
  if (int e = check_mat(image, "lpg")) return e;
  if (int e = check_mat(map, "lpg map")) return e;
  SSK_REQUIRE(map->type == SSK_32FC1 && map->rows == image->rows && map->cols == image->cols, "lpg: map must be CV_32FC1 of the image size");
  SSK_REQUIRE(dscale >= 0 && dscale <= 10 && uscale >= 0 && uscale <= 12, "lpg: dscale 0..10, uscale 0..12");
  SSK_REQUIRE(p >= 0 && p == std::floor(p) && p <= 16, "lpg: only integer powers p are implemented (cv::pow's iPow path)");
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  if (int e = chain_wait(s)) return e;
  Img im;
  if (int e = to_device(image, sc.a, s, &im, 0)) return e;
  float *M = nullptr;
  float *direct = (map->mem == SSK_MEM_DEVICE && map->step == (int64_t)map->cols * 4 && map->data != image->data) ? static_cast<float *>(map->data) : nullptr;
  if (int e = lpg_device(sc, im, k, p, dscale, uscale, &M, direct)) return e;
  if (M != direct) { if (int e = from_device(M, (size_t)im.cols * 4, im.rows, map, s)) return e; }
  return chain_finish(s, on_device(image) && on_device(map));
}

// Host side of c_jovian_derotation_remap / c_saturn_derotation_remap: build_ellipsoid_rotation (ellipsoid.h:47-63,
// build_rotation pose.h:18-58), ellipsoid_bbox (ellipsoid.cc:16-84) and ellipse_crop_box (ellipsoid.cc:279-328), scalar
// double / float arithmetic, no device work.
namespace {
// cv::hypot of lapack.cpp and one Jacobi rotation = cv::eigen of a symmetric 2 x 2 (eigenvalues descending, eigenvectors in rows)
inline double cv_hypot(double a, double b) {
  a = std::fabs(a); b = std::fabs(b);
  if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
  return 0;
}
void eigen_sym2(double a00, double a01, double a11, double W[2], double V[2][2]) {
  W[0] = a00; W[1] = a11;
  V[0][0] = 1; V[0][1] = 0; V[1][0] = 0; V[1][1] = 1;
  const double p = a01;
  if (std::fabs(p) > DBL_EPSILON) {
    const double y = (W[1] - W[0]) * 0.5;
    double t = std::fabs(y) + cv_hypot(p, y);
    double sn = cv_hypot(p, t);
    const double c = t / sn;
    sn = p / sn; t = (p / t) * p;
    if (y < 0) { sn = -sn; t = -t; }
    W[0] -= t; W[1] += t;
    for (int i = 0; i < 2; ++i) {
      const double v0 = V[0][i], v1 = V[1][i];
      V[0][i] = v0 * c - v1 * sn;
      V[1][i] = v0 * sn + v1 * c;
    }
  }
  if (W[0] < W[1]) { std::swap(W[0], W[1]); std::swap(V[0][0], V[1][0]); std::swap(V[0][1], V[1][1]); }
}
void mat3_mul(const double a[9], const double b[9], double c[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
}  // namespace

int ssk_set_stream_ordered(int enable) { return set_stream_ordered(enable); }

int ssk_device_synchronize(void) {
  if (int e = chain_drain()) return e;
  SSK_CUDA(cudaDeviceSynchronize());
  return SSK_OK;
}

int ssk_build_ellipsoid_rotation(const double pose[3], double R[9]) {
  SSK_REQUIRE(pose && R, "build_ellipsoid_rotation: null argument");
  const double ax = pose[1], ay = pose[0], az = pose[2];   // build_rotation(tilt_to_earth, longitude_rotation, position_angle)
  const double cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay), cz = std::cos(az), sz = std::sin(az);
  const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  double t[9];
  mat3_mul(Rz, Rx, t);
  mat3_mul(t, Ry, R);
  return SSK_OK;
}

int ssk_ellipsoid_bbox(int rows, int cols, const double center[2], const double axes[3], const double R[9], float ebox[5], int crop_box[4]) {
  SSK_REQUIRE(center && axes && R && ebox && crop_box, "ellipsoid_bbox: null argument");
  SSK_REQUIRE(rows > 0 && cols > 0 && axes[0] > 0 && axes[1] > 0 && axes[2] > 0, "ellipsoid_bbox: bad geometry");
  // Q = RR diag(1/A^2, 1/B^2, 1/C^2, -1) RR^T is block diagonal, so (P Q^-1 P^T)^-1 restricted to x, y is the inverse of the
  // upper-left 2 x 2 of (R diag(1/A^2, 1/B^2, 1/C^2) R^T)^-1
  const double d[3] = {1 / (axes[0] * axes[0]), 1 / (axes[1] * axes[1]), 1 / (axes[2] * axes[2])};
  double M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = R[i * 3] * d[0] * R[j * 3] + R[i * 3 + 1] * d[1] * R[j * 3 + 1] + R[i * 3 + 2] * d[2] * R[j * 3 + 2];
  const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[2] * M[7] - M[1] * M[8], c11 = M[0] * M[8] - M[2] * M[6];
  const double det = M[0] * c00 + M[1] * (M[5] * M[6] - M[3] * M[8]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
  SSK_REQUIRE(det != 0, "ellipsoid_bbox: singular pose");
  const double s00 = c00 / det, s01 = c01 / det, s11 = c11 / det;   // upper-left 2 x 2 of M^-1
  const double sd = s00 * s11 - s01 * s01;
  SSK_REQUIRE(sd != 0, "ellipsoid_bbox: degenerate outline");
  double W[2], V[2][2];
  eigen_sym2(s11 / sd, -s01 / sd, s00 / sd, W, V);
  const double axis_x = 2 / std::sqrt(W[1]), axis_y = 2 / std::sqrt(W[0]);
  double t0 = std::atan2(V[1][1], V[1][0]);
  if (t0 > CV_PI_D) t0 -= 2 * CV_PI_D;
  if (t0 < -CV_PI_D) t0 += CV_PI_D;
  ebox[0] = (float)center[0]; ebox[1] = (float)center[1]; ebox[2] = (float)axis_x; ebox[3] = (float)axis_y; ebox[4] = (float)(t0 * 180 / CV_PI_D);
  // ellipse_bounding_box: float arithmetic on the RotatedRect members (the angle conversion goes through double)
  const float a = (float)(ebox[4] * CV_PI_D / 180);
  const float ca = std::cos(a), sa = std::sin(a);
  const float ux = ebox[2] * ca / 2, uy = -ebox[2] * sa / 2, vx = ebox[3] * sa / 2, vy = ebox[3] * ca / 2;
  const float hw = std::sqrt(ux * ux + vx * vx), hh = std::sqrt(uy * uy + vy * vy);
  const float left = ebox[0] - hw, top = ebox[1] - hh;
  int x = (int)left, y = (int)top, w = (int)(2 * hw), h = (int)(2 * hh);
  const int margin = 1;                                       // ellipse_crop_box's default, what compute_ellipsoid_zrotation_remap uses
  x -= margin; y -= margin; w += 2 * margin; h += 2 * margin;
  if (x < 0) x = 0;
  if (y < 0) y = 0;
  if (x + w >= cols) w = cols - x;
  if (y + h >= rows) h = rows - y;
  crop_box[0] = x; crop_box[1] = y; crop_box[2] = w; crop_box[3] = h;
  return SSK_OK;
}

// compute_ellipsoid_zrotation_remap (ellipsoid.cc:206-277).  The scalar geometry of the bounding ellipse
// (ellipsoid_bbox / ellipse_crop_box, ellipsoid.cc:16-84, 299-328) stays with the caller.
int ssk_ellipsoid_zrotation_remap(int rows, int cols, const double center[2], const double axes[3], const double R1[9],
                                  const double R2[9], double ebox_angle_deg, const int crop_box[4], double wscale,
                                  ssk_mat *rmap, ssk_mat *wmap, ssk_mat *rmask) {
  if (int e = ensure_device()) return e;
  SSK_REQUIRE(center && axes && R1 && R2 && crop_box, "ellipsoid remap: null argument");
  SSK_REQUIRE(rows > 0 && cols > 0 && axes[0] > 0 && axes[1] > 0 && axes[2] > 0, "ellipsoid remap: bad geometry");
  if (int e = check_mat(rmap, "ellipsoid rmap")) return e;
  if (int e = check_mat(wmap, "ellipsoid wmap")) return e;
  if (int e = check_mat(rmask, "ellipsoid rmask")) return e;
  SSK_REQUIRE(rmap->type == SSK_32FC2 && wmap->type == SSK_32FC1 && rmask->type == SSK_8UC1, "ellipsoid remap: rmap CV_32FC2, wmap CV_32FC1, rmask CV_8UC1");
  SSK_REQUIRE(rmap->rows == rows && rmap->cols == cols && wmap->rows == rows && wmap->cols == cols && rmask->rows == rows && rmask->cols == cols,
              "ellipsoid remap: output size differs from the frame size");
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  Tables tab;
  if (int e = get_tables(&tab)) return e;
  const size_t n = (size_t)rows * cols;
  if (int e = sc.a.ensure(n * 8)) return e;
  if (int e = sc.b.ensure(n * 4)) return e;
  if (int e = sc.c.ensure(n)) return e;
  if (int e = sc.d.ensure(n * 4)) return e;
  EllipsoidArgs a = {};
  a.rows = rows; a.cols = cols; a.cx = center[0]; a.cy = center[1]; a.A = axes[0]; a.B = axes[1]; a.C = axes[2];
  for (int i = 0; i < 9; ++i) { a.R1[i] = R1[i]; a.R2[i] = R2[i]; }
  a.bx = crop_box[0]; a.by = crop_box[1]; a.bw = crop_box[2]; a.bh = crop_box[3];
  const double ang = ebox_angle_deg * 3.1415926535897932384626433832795 / 180;
  a.ca = std::cos(ang); a.sa = std::sin(ang);
  a.wscale = wscale;
  a.rmap = sc.a.as<float2>(); a.wmap = sc.b.as<float>(); a.rmask = sc.c.as<uint8_t>();
  if (int e = launch_ellipsoid_remap(a, s)) return e;
  // cv::remap(wmap, wmap, rmap, INTER_LINEAR, BORDER_CONSTANT)
  RemapArgs ra = {};
  ra.src.data = sc.b.p; ra.src.step = (int64_t)cols * 4; ra.src.rows = rows; ra.src.cols = cols; ra.src.depth = SSK_32F; ra.src.cn = 1; ra.src.scale = 1.f;
  ra.dst = sc.d.as<float>(); ra.dst_step = (int64_t)cols * 4; ra.rows = rows; ra.cols = cols;
  ra.rmap = sc.a.as<float2>(); ra.rmap_step = (int64_t)cols * 8;
  ra.interp = SSK_INTER_LINEAR; ra.border = SSK_BORDER_CONSTANT;
  if (int e = launch_remap(ra, tab, s)) return e;
  if (int e = from_device(sc.a.p, (size_t)cols * 8, rows, rmap, s)) return e;
  if (int e = from_device(sc.d.p, (size_t)cols * 4, rows, wmap, s)) return e;
  if (int e = from_device(sc.c.p, (size_t)cols, rows, rmask, s)) return e;
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

// debayer_nn2(src, dst, colorid) (core/io/debayer.cc:827-1195): raw Bayer frame -> BGR of the same depth
int ssk_debayer_nn2(const ssk_mat *src, ssk_mat *dst, int colorid) {
  if (int e = ensure_device()) return e;
  if (int e = check_mat(src, "debayer_nn2 src")) return e;
  if (int e = check_mat(dst, "debayer_nn2 dst")) return e;
  const int d = type_depth(src->type);
  SSK_REQUIRE(type_cn(src->type) == 1, "debayer_nn2: the Bayer image must have one channel");
  SSK_REQUIRE(type_depth(dst->type) == d && type_cn(dst->type) == 3 && dst->rows == src->rows && dst->cols == src->cols,
              "debayer_nn2: the destination must have 3 channels of the source depth and the source size");
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  Img im;
  if (int e = to_device(src, sc.a, s, &im, 0)) return e;
  const size_t rowb = (size_t)im.cols * 3 * depth_bytes(d);
  void *d_out = dst->data;
  int64_t ostep = dst->step;
  if (dst->mem != SSK_MEM_DEVICE) {
    if (int e = sc.b.ensure(rowb * im.rows)) return e;
    d_out = sc.b.p; ostep = (int64_t)rowb;
  }
  if (int e = launch_debayer_nn2(im.data, im.step, d, im.rows, im.cols, colorid, d_out, ostep, s)) return e;
  if (dst->mem != SSK_MEM_DEVICE) if (int e = from_device(d_out, rowb, im.rows, dst, s)) return e;
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

// unsharp_mask(src, dst, sigma, alpha, outmin, outmax) (core/proc/unsharp_mask.cc:72-118) on CV_32F images: the sharpening of
// the master / reference frame (c_image_stacking_pipeline.cc:1302-1306).  create_lpass_image's single-pass branch
// (unsharp_mask.cc:44-47: sigma <= 2, or an image too small for a pyramid level) handles 1 to 4 channels; its pyrDown /
// pyrUp approximation for larger sigma (unsharp_mask.cc:50-68) single-channel images.
int ssk_unsharp_mask(const ssk_mat *src, ssk_mat *dst, double sigma, double alpha, double outmin, double outmax) {
  if (int e = ensure_device()) return e;
  if (int e = check_mat(src, "unsharp_mask src")) return e;
  if (int e = check_mat(dst, "unsharp_mask dst")) return e;
  SSK_REQUIRE(type_depth(src->type) == SSK_32F && dst->type == src->type && dst->rows == src->rows && dst->cols == src->cols,
              "unsharp_mask: CV_32F source and destination of the same size and type");
  SSK_REQUIRE(alpha < 1.0, "unsharp_mask: alpha < 1");
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  Img im;
  if (int e = to_device(src, sc.a, s, &im, 0)) return e;
  const int cn = im.cn;
  const size_t rowb = (size_t)im.cols * cn * 4, n = (size_t)im.rows * im.cols * cn;
  if (int e = sc.b.ensure(n * 4 * 2)) return e;
  const float *d_src = static_cast<const float *>(im.data);
  if (im.step != (int64_t)rowb) {
    if (int e = sc.c.ensure(n * 4)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(sc.c.p, rowb, im.data, im.step, rowb, im.rows, cudaMemcpyDeviceToDevice, s));
    d_src = sc.c.as<float>();
  }
  float *d_lp = sc.b.as<float>(), *d_out = sc.b.as<float>() + n;
  const int clamp = outmax > outmin;
  if (sigma <= 0 || alpha <= 0) {           // unsharp_mask.cc:76-78: copy (the clamp below still applies)
    if (int e = launch_add_weighted(d_src, 1.0, d_src, 0.0, d_out, (int64_t)n, clamp, (float)outmin, (float)outmax, s)) return e;
  } else {
    int level = 0, Ci = 0;                  // unsharp_mask.cc:25-42
    if (sigma > 2) {
      int m = im.rows < im.cols ? im.rows : im.cols, imax = 0;
      while (m >>= 1) ++imax;
      const int Cc = (int)(sigma * sigma / 2);
      while (level < imax && (1 + 4 * Ci) <= Cc) { Ci = 1 + 4 * Ci; ++level; }
    }
    auto gaussian = [&](const float *from, float *to, int rows, int cols, int ch, double sg) -> int {
      // gaussian_blur of unsharp_mask.cc:19-23: getGaussianKernel(2 * max(1, (int)(5 sigma)) + 1, sigma, CV_32F), BORDER_REFLECT
      SepFilterArgs f = {};
      f.src = from; f.dst = to; f.rows = rows; f.cols = cols; f.batch = 1; f.cn = ch; f.border = SSK_BORDER_REFLECT;
      const int taps = 2 * ((int)(sg * 5) > 1 ? (int)(sg * 5) : 1) + 1;
      SSK_REQUIRE(taps <= kMaxTaps, "unsharp_mask: Gaussian kernel too large");
      double cf[kMaxTaps], sum = 0;
      const double s2 = -0.5 / (sg * sg);
      for (int i = 0; i < taps; ++i) { const double x = i - (taps - 1) * 0.5; cf[i] = std::exp(s2 * x * x); sum += cf[i]; }
      for (int i = 0; i < taps; ++i) f.kx[i] = f.ky[i] = (float)(cf[i] / sum);
      f.kxn = f.kyn = taps;
      return launch_sepfilter(f, s);
    };
    if (level < 1) {
      if (int e = gaussian(d_src, d_lp, im.rows, im.cols, cn, sigma)) return e;
    } else {
      // unsharp_mask.cc:50-68: pyrDown chain (BORDER_REFLECT), residual blur, pyrUp chain back through the size history
      SSK_REQUIRE(cn == 1, "unsharp_mask: the pyramid approximation (sigma > 2) is implemented for single-channel images");
      SSK_REQUIRE(level < 30, "unsharp_mask: too many pyramid levels");
      int lr[32], lc[32];
      lr[0] = im.rows; lc[0] = im.cols;
      for (int j = 1; j <= level; ++j) { lr[j] = (lr[j - 1] + 1) / 2; lc[j] = (lc[j - 1] + 1) / 2; }
      if (int e = sc.d.ensure(n * 4 * 2)) return e;
      float *pp[2] = {sc.d.as<float>(), sc.d.as<float>() + n};
      const float *cur = d_src;
      int w = 0;
      for (int j = 1; j <= level; ++j) {
        PyrDownArgs pd = {};
        pd.src.data = cur; pd.src.step = (int64_t)lc[j - 1] * 4; pd.src.rows = lr[j - 1]; pd.src.cols = lc[j - 1];
        pd.src.depth = SSK_32F; pd.src.cn = 1; pd.src.scale = 1.f;
        pd.dst = pp[w]; pd.dst_rows = lr[j]; pd.dst_cols = lc[j]; pd.batch = 1; pd.post_scale = 1.f; pd.border = SSK_BORDER_REFLECT;
        if (int e = launch_pyrdown(pd, s)) return e;
        cur = pp[w]; w ^= 1;
      }
      const double delta = std::sqrt(sigma * sigma - 2 * Ci) / (double)(1 << level);
      if (delta > 0) {
        if (int e = gaussian(cur, pp[w], lr[level], lc[level], 1, delta)) return e;
        cur = pp[w]; w ^= 1;
      }
      for (int j = level - 1; j >= 0; --j) {
        PyrUpArgs pu = {};
        pu.src = cur; pu.rows = lr[j + 1]; pu.cols = lc[j + 1];
        pu.dst = j == 0 ? d_lp : pp[w]; pu.dst_rows = lr[j]; pu.dst_cols = lc[j]; pu.batch = 1;
        if (int e = launch_pyrup(pu, s)) return e;
        cur = pu.dst; w ^= 1;
      }
    }
    if (int e = launch_add_weighted(d_src, 1.0 / (1.0 - alpha), d_lp, -alpha / (1.0 - alpha), d_out, (int64_t)n, clamp,
                                    (float)outmin, (float)outmax, s)) return e;
  }
  if (int e = from_device(d_out, rowb, im.rows, dst, s)) return e;
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

// average_pyramid_inpaint(src, mask, dst, dstmask, max_levels) (core/proc/inpaint/average_pyramid_inpaint.cc:97-127)
int ssk_average_pyramid_inpaint(const ssk_mat *src, const ssk_mat *mask, ssk_mat *dst, ssk_mat *dstmask, int max_levels) {
  if (int e = ensure_device()) return e;
  if (int e = check_mat(src, "average_pyramid_inpaint src")) return e;
  if (int e = check_mat(dst, "average_pyramid_inpaint dst")) return e;
  SSK_REQUIRE(type_depth(src->type) == SSK_32F && dst->type == src->type && dst->rows == src->rows && dst->cols == src->cols,
              "average_pyramid_inpaint: CV_32F source and destination of the same size and type");
  if (mask) {
    if (int e = check_mat(mask, "average_pyramid_inpaint mask")) return e;
    SSK_REQUIRE(mask->type == SSK_8UC1 && mask->rows == src->rows && mask->cols == src->cols,
                "average_pyramid_inpaint: the mask must be CV_8UC1 of the image size");
  }
  if (dstmask) {
    SSK_REQUIRE(mask, "average_pyramid_inpaint: dstmask requested without a mask");
    if (int e = check_mat(dstmask, "average_pyramid_inpaint dstmask")) return e;
    SSK_REQUIRE(dstmask->type == SSK_8UC1 && dstmask->rows == src->rows && dstmask->cols == src->cols,
                "average_pyramid_inpaint: dstmask must be CV_8UC1 of the image size");
  }
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  Img im;
  if (int e = to_device(src, sc.a, s, &im, 0)) return e;
  const int cn = im.cn;
  const size_t rowb = (size_t)im.cols * cn * 4;
  if (!mask) {     // average_pyramid_inpaint.cc:104-110: nothing to fill
    SSK_CUDA(cudaMemcpy2DAsync(dst->data, dst->step, im.data, im.step, rowb, im.rows,
                               dst->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
    SSK_CUDA(cudaStreamSynchronize(s));
    return SSK_OK;
  }
  Img mim;
  if (int e = to_device(mask, sc.b, s, &mim, 0)) return e;
  const size_t npx = (size_t)im.rows * im.cols;
  if (int e = sc.c.ensure(rowb * im.rows + npx)) return e;
  if (int e = sc.d.ensure(inpaint_work_bytes(im.rows, im.cols, cn, max_levels))) return e;
  float *d_out = sc.c.as<float>();
  uint8_t *d_omask = reinterpret_cast<uint8_t *>(sc.c.as<char>() + rowb * im.rows);
  int full = 0;
  if (int e = launch_average_pyramid_inpaint(static_cast<const float *>(im.data), im.step, static_cast<const uint8_t *>(mim.data), mim.step,
                                             im.rows, im.cols, cn, max_levels, sc.d.p, d_out, d_omask, &full, s)) return e;
  if (int e = from_device(d_out, rowb, im.rows, dst, s)) return e;
  if (dstmask) if (int e = from_device(d_omask, (size_t)im.cols, im.rows, dstmask, s)) return e;
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

// cv::GaussianBlur(src, dst, Size(), sigma_x, sigma_y, BORDER_REPLICATE) on CV_32FC1 (the weight post-processing of
// c_jdr_pipeline.cc:1228): equals cv::sepFilter2D with getGaussianKernel(cvRound(8 sigma + 1) | 1, sigma, CV_32F).
int ssk_gaussian_blur(const ssk_mat *src, double sigma_x, double sigma_y, ssk_mat *dst) {
  if (int e = ensure_device_ordered()) return e;
  if (int e = check_mat(src, "GaussianBlur src")) return e;
  if (int e = check_mat(dst, "GaussianBlur dst")) return e;
  SSK_REQUIRE(src->type == SSK_32FC1 && dst->type == SSK_32FC1 && dst->rows == src->rows && dst->cols == src->cols,
              "GaussianBlur: CV_32FC1 source and destination of the same size");
  if (sigma_y <= 0) sigma_y = sigma_x;
  SSK_REQUIRE(sigma_x > 0 && sigma_x <= 3.5 && sigma_y <= 3.5, "GaussianBlur: 0 < sigma <= 3.5 (kernels up to 31 taps)");
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  if (int e = chain_wait(s)) return e;
  Img im;
  if (int e = to_device(src, sc.a, s, &im, 0)) return e;
  const size_t n = (size_t)im.rows * im.cols;
  if (int e = sc.b.ensure(n * 4 * 2)) return e;
  const float *d_src = static_cast<const float *>(im.data);
  if (im.step != (int64_t)im.cols * 4) {   // sepFilter works on dense images
    SSK_CUDA(cudaMemcpy2DAsync(sc.b.p, (size_t)im.cols * 4, im.data, im.step, (size_t)im.cols * 4, im.rows, cudaMemcpyDeviceToDevice, s));
    d_src = sc.b.as<float>();
  }
  auto taps = [](double sigma, float *k) {
    int n = ((int)std::lrint(sigma * 4 * 2 + 1)) | 1;
    double cf[kMaxTaps], sum = 0;
    const double s2 = -0.5 / (sigma * sigma);
    for (int i = 0; i < n; ++i) { const double x = i - (n - 1) * 0.5; cf[i] = std::exp(s2 * x * x); sum += cf[i]; }
    for (int i = 0; i < n; ++i) k[i] = (float)(cf[i] / sum);
    return n;
  };
  SepFilterArgs f = {};
  // a dense device destination other than the source is written in place of the scratch image
  const bool direct = dst->mem == SSK_MEM_DEVICE && dst->step == (int64_t)dst->cols * 4 && dst->data != src->data;
  f.src = d_src; f.dst = direct ? static_cast<float *>(dst->data) : sc.b.as<float>() + n; f.rows = im.rows; f.cols = im.cols; f.batch = 1;
  f.kxn = taps(sigma_x, f.kx); f.kyn = taps(sigma_y, f.ky);
  if (int e = launch_sepfilter(f, s)) return e;
  if (!direct) { if (int e = from_device(f.dst, (size_t)im.cols * 4, im.rows, dst, s)) return e; }
  return chain_finish(s, on_device(src) && on_device(dst));
}

// One frame of c_jdr_pipeline::derotate_and_average_frames (c_jdr_pipeline.cc:1184-1236): derotation map for the frame's
// longitude offset, per-frame weight (limb weight * time weight [* remapped lpg], master-frame and mask rules),
// GaussianBlur(1, REPLICATE) of the weight, derotation of the frame (INTER_LINEAR, BORDER_TRANSPARENT in place) and
// c_weigthed_average::add(frame, weights) - all on the device.
int ssk_jdr_derotate_and_add(ssk_acc *acc, const ssk_mat *frame, const ssk_mat *mask, const double center[2],
                             const double axes[3], const double R_current[9], const double R_target[9],
                             double ebox_angle_deg, const int crop_box[4], double wscale, int is_master,
                             int enable_weighted_average, double lpg_k, double lpg_p, int lpg_dscale, int lpg_uscale) {
  if (int e = ensure_device_ordered()) return e;
  SSK_REQUIRE(acc && center && axes && R_current && R_target && crop_box, "jdr: null argument");
  if (int e = check_mat(frame, "jdr frame")) return e;
  SSK_REQUIRE(frame->type == SSK_32FC1, "jdr: CV_32FC1 frames (the pipeline's aligned gray frame)");
  SSK_REQUIRE(!enable_weighted_average || (lpg_p >= 0 && lpg_p == std::floor(lpg_p) && lpg_p <= 16), "jdr: integer lpg power");
  const int rows = frame->rows, cols = frame->cols;
  Scratch &sc = scratch();
  if (int e = sc.init()) return e;
  cudaStream_t s = sc.stream;
  if (int e = chain_wait(s)) return e;
  static thread_local DevBuf d_frame, d_mask, d_rmap, d_wpre, d_w, d_rmask, d_wblur;
  Tables tab;
  if (int e = get_tables(&tab)) return e;
  const size_t n = (size_t)rows * cols;
  if (int e = d_rmap.ensure(n * 8)) return e;
  if (int e = d_wpre.ensure(n * 4)) return e;
  if (int e = d_w.ensure(n * 4)) return e;
  if (int e = d_rmask.ensure(n)) return e;
  if (int e = d_wblur.ensure(n * 4)) return e;
  // the frame as a dense device image (a dense device frame is used where it lies)
  const float *d_f;
  if (frame->mem == SSK_MEM_DEVICE && frame->step == (int64_t)cols * 4) d_f = static_cast<const float *>(frame->data);
  else {
    if (int e = d_frame.ensure(n * 4)) return e;
    SSK_CUDA(cudaMemcpy2DAsync(d_frame.p, (size_t)cols * 4, frame->data, frame->step, (size_t)cols * 4, rows,
                               frame->mem == SSK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    d_f = d_frame.as<float>();
  }
  const uint8_t *dm = nullptr;
  int64_t mstep = 0;
  if (mask) { if (int e = mask_to_device(mask, rows, cols, d_mask, s, &dm, &mstep)) return e; }
  // derotation map, disk mask, limb weight
  EllipsoidArgs a = {};
  a.rows = rows; a.cols = cols; a.cx = center[0]; a.cy = center[1]; a.A = axes[0]; a.B = axes[1]; a.C = axes[2];
  for (int i = 0; i < 9; ++i) { a.R1[i] = R_current[i]; a.R2[i] = R_target[i]; }
  a.bx = crop_box[0]; a.by = crop_box[1]; a.bw = crop_box[2]; a.bh = crop_box[3];
  const double ang = ebox_angle_deg * 3.1415926535897932384626433832795 / 180;
  a.ca = std::cos(ang); a.sa = std::sin(ang); a.wscale = wscale;
  a.rmap = d_rmap.as<float2>(); a.wmap = d_wpre.as<float>(); a.rmask = d_rmask.as<uint8_t>();
  if (int e = launch_ellipsoid_remap(a, s)) return e;
  // lpg(frame); its TRANSPARENT remap, the remap of the limb weight and the weight rules are one pass (k_jdr_weights_fused)
  float *L = nullptr;
  if (enable_weighted_average) {
    Img im;
    im.data = d_f; im.step = (int64_t)cols * 4; im.rows = rows; im.cols = cols; im.depth = SSK_32F; im.cn = 1; im.scale = 1.f;
    if (int e = lpg_device(sc, im, lpg_k, lpg_p, lpg_dscale, lpg_uscale, &L)) return e;
  }
  if (int e = launch_jdr_weights_fused(d_wpre.as<float>(), L, d_rmap.as<float2>(), d_rmask.as<uint8_t>(), dm, mstep, rows, cols, is_master,
                                       d_w.as<float>(), s)) return e;
  // cv::GaussianBlur(weights, Size(), 1, 1, BORDER_REPLICATE): 9 taps
  SepFilterArgs f = {};
  f.src = d_w.as<float>(); f.dst = d_wblur.as<float>(); f.rows = rows; f.cols = cols; f.batch = 1;
  {
    double cf[9], sum = 0;
    for (int i = 0; i < 9; ++i) { const double x = i - 4.0; cf[i] = std::exp(-0.5 * x * x); sum += cf[i]; }
    for (int i = 0; i < 9; ++i) f.kx[i] = f.ky[i] = (float)(cf[i] / sum);
    f.kxn = f.kyn = 9;
  }
  if (int e = launch_sepfilter(f, s)) return e;
  // derotation of the frame (INTER_LINEAR, BORDER_TRANSPARENT in place) fused with c_weigthed_average::add(frame, weights)
  Acc &A = acc->a;
  SSK_REQUIRE(A.kind == SSK_ACC_WEIGHTED_AVERAGE, "jdr: a c_weigthed_average accumulator is expected");
  if (A.rows) SSK_REQUIRE(A.rows == rows && A.cols == cols && A.cn == 1, "c_weigthed_average::add: frame size / channels differ from the accumulator");
  if (int e = A.ensure(rows, cols, 1)) return e;
  if (int e = stream_after(s, A.stream)) return e;     // work the accumulator has in flight on its own stream, the zero-fill of a first use
  if (int e = launch_jdr_remap_add(d_f, d_rmap.as<float2>(), d_wblur.as<float>(), rows, cols, A.acc.as<float>(), A.wacc.as<float>(), s)) return e;
  ++A.frames;
  if (int e = stream_after(A.stream, s)) return e;     // later calls on the accumulator's stream see this frame
  return chain_finish(s, on_device(frame) && on_device(mask));
}

}  // extern "C"
