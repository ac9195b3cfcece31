// ssk_adapter.h - header-only C++ adapter with the reference's class names over the C ABI (include/ssk.h).
//
// A SerStacker maintainer swaps the CPU classes of the stacking hot path for these by including this header
// and linking libssk.so.  Method names, argument meaning and the bool-return error convention
// follow the reference declarations cited at each class; the last error text is ssk_last_error() where the
// reference logs with CF_ERROR.
//
// Image type: with -DSSK_WITH_OPENCV the adapter takes cv::Mat / cv::InputArray like the reference; without
// OpenCV headers (this repository's build image has none) it uses ssk::Mat, a minimal owning matrix with the
// cv::Mat memory layout, so the adapter and its test build with g++ alone.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ssk.h"

#ifdef SSK_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace ssk {

// ---------------------------------------------------------------------------------------------------------
// Minimal matrix (cv::Mat layout: row-major, interleaved channels, OpenCV type code).
// ---------------------------------------------------------------------------------------------------------
struct Mat {
  int rows = 0, cols = 0, type = SSK_32FC1;
  std::vector<uint8_t> buf;

  Mat() = default;
  Mat(int r, int c, int t) { create(r, c, t); }
  static int elem_size(int t) {
    const int depth = t & 7, cn = (t >> 3) + 1;
    return (depth == SSK_8U ? 1 : depth == SSK_16U ? 2 : 4) * cn;
  }
  void create(int r, int c, int t) {
    rows = r, cols = c, type = t;
    buf.assign((size_t)r * c * elem_size(t), 0);
  }
  bool empty() const { return buf.empty(); }
  int channels() const { return (type >> 3) + 1; }
  size_t step() const { return (size_t)cols * elem_size(type); }
  template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(buf.data() + r * step()); }
  template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(buf.data() + r * step()); }
};

namespace detail {
inline ssk_mat view(const Mat &m) {
  ssk_mat v;
  v.data = const_cast<uint8_t *>(m.buf.data());
  v.step = (int64_t)m.step();
  v.rows = m.rows, v.cols = m.cols, v.type = m.type, v.mem = SSK_MEM_HOST;
  return v;
}
#ifdef SSK_WITH_OPENCV
inline ssk_mat view(const cv::Mat &m) {
  ssk_mat v;
  v.data = m.data;
  v.step = (int64_t)m.step;
  v.rows = m.rows, v.cols = m.cols, v.type = m.type(), v.mem = SSK_MEM_HOST;
  return v;
}
#endif
// optional argument: empty matrix <=> cv::noArray()
template <class M> struct Opt {
  ssk_mat v;
  bool have;
  explicit Opt(const M &m) : v(view(m)), have(!m.empty()) {}
  const ssk_mat *get() const { return have ? &v : nullptr; }
  ssk_mat *get() { return have ? &v : nullptr; }
};
}  // namespace detail

#ifdef SSK_WITH_OPENCV
using image_t = cv::Mat;
inline void create_like(image_t &m, int rows, int cols, int type) { m.create(rows, cols, type); }
#else
using image_t = Mat;
inline void create_like(image_t &m, int rows, int cols, int type) { m.create(rows, cols, type); }
#endif

// ---------------------------------------------------------------------------------------------------------
// c_image_transform and create_image_transform()  (core/proc/image_registration/c_image_transform.h:42-117,
// image_transform.cc:34-63).  One value class instead of the reference's five subclasses: the kind is
// motion_type(), the parameter vector has the reference's layout.
// ---------------------------------------------------------------------------------------------------------
class c_image_transform {
 public:
  typedef std::shared_ptr<c_image_transform> sptr;
  explicit c_image_transform(int motion_type) { ssk_transform_init(&t_, motion_type); }
  int motion_type() const { return t_.motion_type; }
  std::vector<float> parameters() const { return std::vector<float>(t_.params, t_.params + t_.nparams); }
  bool set_parameters(const std::vector<float> &p) {
    if ((int)p.size() != t_.nparams) return false;
    std::memcpy(t_.params, p.data(), sizeof(float) * p.size());
    return true;
  }
  void scale_transfrom(double factor) { ssk_transform_scale(&t_, factor); }   // (sic) reference spelling
  bool create_remap(int cols, int rows, image_t &rmap) const {
    create_like(rmap, rows, cols, SSK_32FC2);
    ssk_mat v = detail::view(rmap);
    return ssk_transform_create_remap(&t_, rows, cols, &v) == SSK_OK;
  }
  ssk_transform &raw() { return t_; }
  const ssk_transform &raw() const { return t_; }

 private:
  ssk_transform t_;
};

inline c_image_transform::sptr create_image_transform(int motion_type) {
  return std::make_shared<c_image_transform>(motion_type);
}

// ---------------------------------------------------------------------------------------------------------
// c_ecch  (core/proc/image_registration/ecc2.h:193-308)
// ---------------------------------------------------------------------------------------------------------
class c_ecch {
 public:
  explicit c_ecch(c_image_transform *transform = nullptr, int method = SSK_ECC_INVERSE_COMPOSITIONAL_LM)
      : transform_(transform) {
    ssk_ecch_options_default(&opts_);
    opts_.method = method;
  }
  ~c_ecch() { if (h_) ssk_ecch_destroy(h_); }
  c_ecch(const c_ecch &) = delete;
  c_ecch &operator=(const c_ecch &) = delete;

  const ssk_ecch_options &options() const { return opts_; }
  void set_options(const ssk_ecch_options &o) { opts_ = o; reset(); }
  void set_image_transform(c_image_transform *t) { transform_ = t; }
  c_image_transform *image_transform() const { return transform_; }
  void set_method(int v) { opts_.method = v; reset(); }
  void set_maxlevel(int v) { opts_.maxlevel = v; reset(); }
  void set_minimum_image_size(int v) { opts_.minimum_image_size = v; reset(); }
  void set_max_iterations(int v) { opts_.max_iterations = v; reset(); }
  void set_epsx(double v) { opts_.epsx = v; reset(); }
  void set_interpolation(int v) { opts_.interpolation = v; reset(); }
  void set_input_smooth_sigma(double v) { opts_.input_smooth_sigma = v; reset(); }
  void set_reference_smooth_sigma(double v) { opts_.reference_smooth_sigma = v; reset(); }
  void set_update_step_scale(double v) { opts_.update_step_scale = v; reset(); }

  bool set_reference_image(const image_t &reference_image, const image_t &reference_mask = image_t()) {
    if (!h_ && ssk_ecch_create(&opts_, &h_) != SSK_OK) return false;
    ssk_mat im = detail::view(reference_image);
    detail::Opt<image_t> mk(reference_mask);
    return ssk_ecch_set_reference_image(h_, &im, mk.get()) == SSK_OK;
  }
  bool align(const image_t &current_image, const image_t &current_mask = image_t()) {
    if (!h_ || !transform_) return false;
    ssk_mat im = detail::view(current_image);
    detail::Opt<image_t> mk(current_mask);
    return ssk_ecch_align(h_, &im, mk.get(), &transform_->raw(), &status_) == SSK_OK && !status_.failed;
  }
  double eps() const { return status_.eps; }
  int num_iterations() const { return status_.num_iterations; }
  bool failed() const { return status_.failed != 0; }

 private:
  void reset() { if (h_) { ssk_ecch_destroy(h_); h_ = nullptr; } }   // options are bound at creation
  ssk_ecch *h_ = nullptr;
  ssk_ecch_options opts_;
  ssk_ecc_status status_ = {};
  c_image_transform *transform_;
};

// ---------------------------------------------------------------------------------------------------------
// c_frame_registration, ECC branch  (core/proc/image_registration/c_frame_registration.h:210-354)
// ---------------------------------------------------------------------------------------------------------
struct c_image_registration_status { ssk_ecc_status ecc = {}; };

class c_frame_registration {
 public:
  typedef std::shared_ptr<c_frame_registration> sptr;
  c_frame_registration() { ssk_registration_options_default(&opts_); }
  explicit c_frame_registration(const ssk_registration_options &o) : opts_(o) {}
  ~c_frame_registration() { if (h_) ssk_reg_destroy(h_); }
  c_frame_registration(const c_frame_registration &) = delete;
  c_frame_registration &operator=(const c_frame_registration &) = delete;

  ssk_registration_options &options() { return opts_; }
  const c_image_transform::sptr &image_transform() const { return transform_; }
  const c_image_registration_status &status() const { return status_; }
  void set_input_bpp(int bpp) { bpp_ = bpp; }   // 8U/16U frames: c_image_stacking_pipeline_base.cc:271-276 scaling

  bool setup_reference_frame(const image_t &image, const image_t &msk = image_t()) {
    if (h_) { ssk_reg_destroy(h_); h_ = nullptr; }
    if (ssk_reg_create(&opts_, &h_) != SSK_OK) return false;
    ssk_mat im = detail::view(image);
    detail::Opt<image_t> mk(msk);
    return ssk_reg_setup_reference_frame(h_, &im, mk.get(), bpp_) == SSK_OK;
  }
  // register_frame(src, srcmask, dst, dstmask): estimates the transform; dst/dstmask non-null => remap too.
  bool register_frame(const image_t &src, const image_t &srcmask = image_t(), image_t *dst = nullptr,
                      image_t *dstmask = nullptr) {
    if (!h_) return false;
    if (!transform_) transform_ = create_image_transform(opts_.motion_type);
    ssk_mat im = detail::view(src);
    detail::Opt<image_t> mk(srcmask);
    if (ssk_reg_register_frame(h_, &im, mk.get(), bpp_, &transform_->raw(), &status_.ecc) != SSK_OK) return false;
    return dst ? remap(src, *dst, srcmask, dstmask) : true;
  }
  bool remap(const image_t &src, image_t &dst, const image_t &src_mask = image_t(), image_t *dst_mask = nullptr,
             int interpolation = -1, int border_mode = -1, const double border_value[4] = nullptr) const {
    return custom_remap(image_t(), src, dst, src_mask, dst_mask, interpolation, border_mode, border_value);
  }
  bool custom_remap(const image_t &rmap, const image_t &src, image_t &dst, const image_t &src_mask = image_t(),
                    image_t *dst_mask = nullptr, int interpolation = -1, int border_mode = -1,
                    const double border_value[4] = nullptr) const {
    if (!h_) return false;
    ssk_mat s = detail::view(src);
    int cols = 0, rows = 0;
    if (!rmap.empty()) rows = detail::view(rmap).rows, cols = detail::view(rmap).cols;
    else rows = s.rows, cols = s.cols;
    create_like(dst, rows, cols, SSK_MAKETYPE(SSK_32F, (s.type >> 3) + 1));
    ssk_mat d = detail::view(dst), dm;
    if (dst_mask) { create_like(*dst_mask, rows, cols, SSK_8UC1); dm = detail::view(*dst_mask); }
    detail::Opt<image_t> rm(rmap), sm(src_mask);
    static const double zero[4] = {0, 0, 0, 0};
    return ssk_reg_remap(h_, rm.get(), &s, &d, sm.get(), dst_mask ? &dm : nullptr, interpolation, border_mode,
                         border_value ? border_value : zero) == SSK_OK;
  }
  bool current_remap(image_t &rmap, int cols, int rows) const {
    if (!h_) return false;
    create_like(rmap, rows, cols, SSK_32FC2);
    ssk_mat v = detail::view(rmap);
    return ssk_reg_get_current_remap(h_, &v) == SSK_OK;
  }

 private:
  ssk_reg *h_ = nullptr;
  ssk_registration_options opts_;
  c_image_transform::sptr transform_;
  c_image_registration_status status_;
  int bpp_ = 0;
};

// ---------------------------------------------------------------------------------------------------------
// c_frame_accumulation / c_weigthed_average / c_bayer_average  (core/average/c_frame_accumulation.h:14-63, 222-262)
// ---------------------------------------------------------------------------------------------------------
class c_frame_accumulation {
 public:
  typedef std::shared_ptr<c_frame_accumulation> ptr;
  virtual ~c_frame_accumulation() { if (h_) ssk_acc_destroy(h_); }
  bool add(const image_t &src, const image_t &mask_or_weights = image_t()) {
    ssk_mat s = detail::view(src);
    detail::Opt<image_t> w(mask_or_weights);
    return ssk_acc_add(h_, &s, w.get(), 0) == SSK_OK;
  }
  bool compute(image_t &avg, image_t *mask = nullptr, double dscale = 1.0) const {
    int cols = 0, rows = 0, cn = 0;
    if (ssk_acc_size(h_, &cols, &rows, &cn) != SSK_OK || cols <= 0) return false;
    create_like(avg, rows, cols, SSK_MAKETYPE(SSK_32F, cn));
    ssk_mat a = detail::view(avg), m;
    if (mask) { create_like(*mask, rows, cols, SSK_8UC1); m = detail::view(*mask); }
    return ssk_acc_compute(h_, &a, mask ? &m : nullptr, dscale) == SSK_OK;
  }
  // compute() + average_pyramid_inpaint(avg, mask, avg, mask, max_levels) on the device (c_image_stacking_pipeline.cc:742-767)
  bool compute_inpainted(image_t &avg, image_t *mask = nullptr, double dscale = 1.0, int max_levels = 100) const {
    int cols = 0, rows = 0, cn = 0;
    if (ssk_acc_size(h_, &cols, &rows, &cn) != SSK_OK || cols <= 0) return false;
    create_like(avg, rows, cols, SSK_MAKETYPE(SSK_32F, cn));
    ssk_mat a = detail::view(avg), m;
    if (mask) { create_like(*mask, rows, cols, SSK_8UC1); m = detail::view(*mask); }
    return ssk_acc_compute_inpainted(h_, &a, mask ? &m : nullptr, dscale, max_levels) == SSK_OK;
  }
  bool reinitialize(const image_t &src, const image_t &accw) {
    ssk_mat s = detail::view(src), w = detail::view(accw);
    return ssk_acc_reinitialize(h_, &s, &w) == SSK_OK;
  }
  void clear() { ssk_acc_clear(h_); }
  int accumulated_frames() const { return ssk_acc_frames(h_); }
  bool accumulator_size(int *cols, int *rows) const { int cn; return ssk_acc_size(h_, cols, rows, &cn) == SSK_OK; }
  ssk_acc *handle() const { return h_; }

 protected:
  explicit c_frame_accumulation(int kind) { ssk_acc_create(kind, &h_); }
  ssk_acc *h_ = nullptr;
};

class c_weigthed_average : public c_frame_accumulation {   // (sic) reference spelling
 public:
  c_weigthed_average() : c_frame_accumulation(SSK_ACC_WEIGHTED_AVERAGE) {}
};

class c_bayer_average : public c_frame_accumulation {
 public:
  c_bayer_average() : c_frame_accumulation(SSK_ACC_BAYER_AVERAGE) {}
  void set_bayer_pattern(int colorid) { ssk_acc_set_bayer_pattern(h_, colorid); }
  bool set_remap(const image_t &rmap) {
    ssk_mat v = detail::view(rmap);
    return ssk_acc_set_remap(h_, nullptr, &v) == SSK_OK;
  }
};

// c_local_variance_sharpness_measure::compute (c_local_variance_sharpness_measure.cc:193-247)
inline bool compute_local_variance_map(const image_t &image, image_t &map, int dscale = 1, int kradius = 1,
                                       int uscale = 0, double *Q = nullptr) {
  ssk_mat s = detail::view(image);
  create_like(map, s.rows, s.cols, SSK_32FC1);
  ssk_mat m = detail::view(map);
  double q = 0;
  const bool ok = ssk_local_variance_map(&s, 0, dscale, kradius, uscale, &m, &q) == SSK_OK;
  if (Q) *Q = q;
  return ok;
}

// lpg (core/proc/lpg.cc:223-290): Laplacian + gradient energy weight map (integer powers p)
inline bool lpg(const image_t &image, image_t &map, double k = 2.0, double p = 2.0, int dscale = 2, int uscale = 6) {
  ssk_mat s = detail::view(image);
  create_like(map, s.rows, s.cols, SSK_32FC1);
  ssk_mat m = detail::view(map);
  return ssk_lpg(&s, k, p, dscale, uscale, &m) == SSK_OK;
}

// debayer_nn2 (core/io/debayer.cc:827-1195): raw Bayer frame -> BGR of the same depth; colorid = SSK_COLORID_BAYER_*
inline bool debayer_nn2(const image_t &src, image_t &dst, int colorid) {
  ssk_mat s = detail::view(src);
  image_t out;
  create_like(out, s.rows, s.cols, SSK_MAKETYPE(s.type & 7, 3));
  ssk_mat d = detail::view(out);
  const bool ok = ssk_debayer_nn2(&s, &d, colorid) == SSK_OK;
  if (ok) dst = out;
  return ok;
}

// unsharp_mask (core/proc/unsharp_mask.cc:72-118): sharpening of the master / reference frame
// (c_image_stacking_pipeline.cc:1302-1306).  CV_32F; outmax <= outmin: no clamp.
inline bool unsharp_mask(const image_t &src, image_t &dst, double sigma, double alpha, double outmin = -1, double outmax = -1) {
  ssk_mat s = detail::view(src);
  image_t out;
  create_like(out, s.rows, s.cols, s.type);
  ssk_mat d = detail::view(out);
  const bool ok = ssk_unsharp_mask(&s, &d, sigma, alpha, outmin, outmax) == SSK_OK;
  if (ok) dst = out;      // src and dst may be the same image, as at the reference's call site
  return ok;
}

// average_pyramid_inpaint (core/proc/inpaint/average_pyramid_inpaint.cc:97-127; call site
// c_image_stacking_pipeline.cc:763-767 with max_levels = 100).  src CV_32F, mask CV_8UC1.
inline bool average_pyramid_inpaint(const image_t &src, const image_t &mask, image_t &dst, image_t *dstmask = nullptr,
                                    int max_levels = 100) {
  ssk_mat s = detail::view(src), m = detail::view(mask);
  image_t out, outmask;
  create_like(out, s.rows, s.cols, s.type);
  ssk_mat d = detail::view(out), dm;
  if (dstmask) { create_like(outmask, s.rows, s.cols, SSK_8UC1); dm = detail::view(outmask); }
  const bool ok = ssk_average_pyramid_inpaint(&s, &m, &d, dstmask ? &dm : nullptr, max_levels) == SSK_OK;
  if (ok) { dst = out; if (dstmask) *dstmask = outmask; }
  return ok;
}

// compute_ellipsoid_zrotation_remap (core/proc/feature2d/ellipsoid.cc:206-277).  R1 / R2: row-major 3x3 doubles;
// ebox_angle_deg / crop_box {x, y, w, h}: ellipsoid_bbox(center, A, B, C, R2).angle and ellipse_crop_box(ebox, size).
inline bool compute_ellipsoid_zrotation_remap(int rows, int cols, const double center[2], const double axes[3],
                                              const double R1[9], const double R2[9], double ebox_angle_deg,
                                              const int crop_box[4], double wscale, image_t &rmap, image_t &wmap,
                                              image_t &rmask) {
  create_like(rmap, rows, cols, SSK_32FC2);
  create_like(wmap, rows, cols, SSK_32FC1);
  create_like(rmask, rows, cols, SSK_8UC1);
  ssk_mat a = detail::view(rmap), b = detail::view(wmap), c = detail::view(rmask);
  return ssk_ellipsoid_zrotation_remap(rows, cols, center, axes, R1, R2, ebox_angle_deg, crop_box, wscale, &a, &b, &c) == SSK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// The batched per-frame loop of c_image_stacking_pipeline::process_input_sequence
// (c_image_stacking_pipeline.cc:1358-1862): one call registers, warps and accumulates a batch of frames.
// ---------------------------------------------------------------------------------------------------------
class c_stacking_loop {
 public:
  explicit c_stacking_loop(const ssk_stack_options &o) { ssk_stack_create(&o, &h_); }
  ~c_stacking_loop() { if (h_) ssk_stack_destroy(h_); }
  c_stacking_loop(const c_stacking_loop &) = delete;
  c_stacking_loop &operator=(const c_stacking_loop &) = delete;
  bool valid() const { return h_ != nullptr; }
  bool set_reference(const image_t &image, int bpp = 0) {
    ssk_mat v = detail::view(image);
    return ssk_stack_set_reference(h_, &v, nullptr, bpp) == SSK_OK;
  }
  bool add_frames(const std::vector<image_t> &frames, int bpp = 0, std::vector<ssk_transform> *transforms = nullptr,
                  std::vector<ssk_ecc_status> *status = nullptr) {
    std::vector<ssk_mat> v;
    for (const image_t &f : frames) v.push_back(detail::view(f));
    if (transforms) transforms->resize(v.size());
    if (status) status->resize(v.size());
    return ssk_stack_add_frames(h_, v.data(), (int)v.size(), bpp, transforms ? transforms->data() : nullptr,
                                status ? status->data() : nullptr) == SSK_OK;
  }
  bool compute(image_t &avg, image_t &mask, int rows, int cols, int cn = 1) {
    create_like(avg, rows, cols, SSK_MAKETYPE(SSK_32F, cn));
    create_like(mask, rows, cols, SSK_8UC1);
    ssk_mat a = detail::view(avg), m = detail::view(mask);
    return ssk_stack_compute(h_, &a, &m) == SSK_OK;
  }
  bool compute_inpainted(image_t &avg, image_t &mask, int rows, int cols, int cn = 1, int max_levels = 100) {
    create_like(avg, rows, cols, SSK_MAKETYPE(SSK_32F, cn));
    create_like(mask, rows, cols, SSK_8UC1);
    ssk_mat a = detail::view(avg), m = detail::view(mask);
    return ssk_stack_compute_inpainted(h_, &a, &m, max_levels) == SSK_OK;
  }
  int accumulated_frames() const { return ssk_stack_accumulated_frames(h_); }
  // streaming form: enqueue a chunk (<= max_batch frames) and collect its per-frame results one chunk late
  bool submit(const std::vector<image_t> &frames, int64_t *ticket, int bpp = 0) {
    std::vector<ssk_mat> v;
    for (const image_t &f : frames) v.push_back(detail::view(f));
    return ssk_stack_submit(h_, v.data(), (int)v.size(), bpp, ticket) == SSK_OK;
  }
  bool wait(int64_t ticket, std::vector<ssk_transform> *transforms, std::vector<ssk_ecc_status> *status, int capacity) {
    if (transforms) transforms->resize(capacity);
    if (status) status->resize(capacity);
    int n = 0;
    const bool ok = ssk_stack_wait(h_, ticket, transforms ? transforms->data() : nullptr, status ? status->data() : nullptr, capacity, &n) == SSK_OK;
    if (transforms) transforms->resize(n);
    if (status) status->resize(n);
    return ok;
  }
  bool sync() { return ssk_stack_sync(h_) == SSK_OK; }
  bool flush() { return ssk_stack_flush(h_) == SSK_OK; }   // stream-side join of the side-stream ring kernel (see ssk.h)

 private:
  ssk_stack *h_ = nullptr;
};

}  // namespace ssk
