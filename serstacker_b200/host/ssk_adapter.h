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
#include <cmath>
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
  void reset() { const int mt = t_.motion_type; ssk_transform_init(&t_, mt); }   // c_image_transform::reset(): identity
  bool invertible() const { return true; }                                    // every transform of the path is (c_image_transform.h:100)
  // translation() / set_translation() (c_image_transform.h:95-98): where the shift lives in each parameter layout
  void translation(float *tx, float *ty) const {
    if (t_.motion_type == SSK_MOTION_AFFINE) { *tx = t_.params[2]; *ty = t_.params[5]; }
    else if (t_.motion_type == SSK_MOTION_HOMOGRAPHY) { *tx = t_.params[2] / t_.aux[2]; *ty = t_.params[5] / t_.aux[2]; }
    else { *tx = t_.params[0]; *ty = t_.params[1]; }
  }
  void set_translation(float tx, float ty) {
    if (t_.motion_type == SSK_MOTION_AFFINE) { t_.params[2] = tx; t_.params[5] = ty; }
    else if (t_.motion_type == SSK_MOTION_HOMOGRAPHY) { t_.params[2] = tx * t_.aux[2]; t_.params[5] = ty * t_.aux[2]; }
    else { t_.params[0] = tx; t_.params[1] = ty; }
  }
  // eps(dp, image_size) and invert_and_compose(parameters(), dp) (c_image_transform.h:52, 107-110): host arithmetic in libssk
  double eps(const std::vector<float> &dp, int cols, int rows) const {
    double e = 0;
    if (ssk_transform_eps(&t_, dp.data(), (int)dp.size(), rows, cols, &e) != SSK_OK) return -1;
    return e;
  }
  std::vector<float> invert_and_compose(const std::vector<float> &dp) const {
    std::vector<float> out((size_t)t_.nparams);
    if (ssk_transform_invert_and_compose(&t_, dp.data(), (int)dp.size(), out.data()) != SSK_OK) out.clear();
    return out;
  }
  // remap(rpts, cpts) (c_image_transform.h:79-82): interleaved (x, y) floats
  bool remap(const std::vector<float> &rpts_xy, std::vector<float> &cpts_xy) const {
    cpts_xy.resize(rpts_xy.size());
    return ssk_transform_remap_points(&t_, rpts_xy.data(), (int)(rpts_xy.size() / 2), cpts_xy.data()) == SSK_OK;
  }
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
  // set_current_image(image, mask) + align(): the two-step form (ecc2.h:262-266); the image is kept until align()
  bool set_current_image(const image_t &current_image, const image_t &current_mask = image_t()) {
    cur_ = current_image; cur_mask_ = current_mask;
    return !cur_.empty();
  }
  bool align() { return !cur_.empty() && align(cur_, cur_mask_); }
  // align(current_image, current_mask, reference_image, reference_mask) (ecc2.h:270-271)
  bool align(const image_t &current_image, const image_t &current_mask, const image_t &reference_image, const image_t &reference_mask) {
    return set_reference_image(reference_image, reference_mask) && align(current_image, current_mask);
  }
  // create_remap(rmap) (ecc2.h:287): the dense CV_32FC2 map of the bound transform at the reference image size
  bool create_remap(image_t &rmap) const {
    int cols = 0, rows = 0;
    if (!h_ || !transform_ || ssk_ecch_level_size(h_, 0, &cols, &rows) != SSK_OK) return false;
    return transform_->create_remap(cols, rows, rmap);
  }
  // reference_image() / current_image() (ecc2.h:275-282): level 0 of the pyramids, as the solver sees them (smoothed)
  bool reference_image(image_t &dst, int level = 0) const { return get_image(0, level, dst); }
  bool current_image(image_t &dst, int level = 0) const { return get_image(1, level, dst); }
  double eps() const { return status_.eps; }
  int num_iterations() const { return status_.num_iterations; }
  bool failed() const { return status_.failed != 0; }
  // static c_ecch::compute_next_pyramid_layer_size (ecc2.h:290-293)
  static void compute_next_pyramid_layer_size(int cols, int rows, int *ncols, int *nrows) {
    *ncols = ((cols + 1) >> 1) & ~1; *nrows = ((rows + 1) >> 1) & ~1;
  }

 private:
  bool get_image(int which, int level, image_t &dst) const {
    int cols = 0, rows = 0;
    if (!h_ || ssk_ecch_level_size(h_, level, &cols, &rows) != SSK_OK) return false;
    create_like(dst, rows, cols, SSK_32FC1);
    ssk_mat v = detail::view(dst);
    return ssk_ecch_get_image(h_, which, level, &v) == SSK_OK;
  }
  image_t cur_, cur_mask_;
  void reset() { if (h_) { ssk_ecch_destroy(h_); h_ = nullptr; } }   // options are bound at creation
  ssk_ecch *h_ = nullptr;
  ssk_ecch_options opts_;
  ssk_ecc_status status_ = {};
  c_image_transform *transform_;
};

// ---------------------------------------------------------------------------------------------------------
// c_eccflow  (core/proc/image_registration/ecc2.h:515-662): dense smooth optical flow
// ---------------------------------------------------------------------------------------------------------
struct c_eccflow_options {   // ecc2.h:515-527
  double input_smooth_sigma = 0;
  double reference_smooth_sigma = 0;
  double update_multiplier = 1.5;
  double scale_factor = 0.5;
  double noise_level = -1;
  int max_iterations = 1;
  int support_scale = 5;
  int min_image_size = 4;
  int max_pyramid_level = -1;
  int downscale = SSK_ECCFLOW_DOWNSCALE_RECURSIVE_RESIZE;
};

class c_eccflow {
 public:
  c_eccflow() = default;
  ~c_eccflow() { if (h_) ssk_eccflow_destroy(h_); }
  c_eccflow(const c_eccflow &) = delete;
  c_eccflow &operator=(const c_eccflow &) = delete;
  c_eccflow_options &options() { return opts_; }
  const c_eccflow_options &options() const { return opts_; }
  void set_options(const c_eccflow_options &o) { opts_ = o; }
  void set_support_scale(int v) { opts_.support_scale = v; }
  int support_scale() const { return opts_.support_scale; }
  void set_max_iterations(int v) { opts_.max_iterations = v; }
  int max_iterations() const { return opts_.max_iterations; }
  void set_update_multiplier(double v) { opts_.update_multiplier = v; }
  double update_multiplier() const { return opts_.update_multiplier; }
  void set_input_smooth_sigma(double v) { opts_.input_smooth_sigma = v; }
  void set_reference_smooth_sigma(double v) { opts_.reference_smooth_sigma = v; }
  void set_downscale_method(int v) { opts_.downscale = v; }
  int downscale_method() const { return opts_.downscale; }
  void set_scale_factor(double v) { opts_.scale_factor = v; }
  double scale_factor() const { return opts_.scale_factor; }
  void set_min_image_size(int v) { opts_.min_image_size = v; }
  int min_image_size() const { return opts_.min_image_size; }
  void set_max_pyramid_level(int v) { opts_.max_pyramid_level = v; }
  int max_pyramid_level() const { return opts_.max_pyramid_level; }
  void set_noise_level(double v) { opts_.noise_level = v; }
  double noise_level() const { return opts_.noise_level; }
  void copy_parameters(const c_eccflow &rhs) { opts_ = rhs.opts_; }

  // the options are latched here, as the reference builds its pyramid here (ecc2.cc:2494-2672)
  bool set_reference_image(const image_t &reference_image, const image_t &reference_mask = image_t()) {
    if (h_) { ssk_eccflow_destroy(h_); h_ = nullptr; }
    ssk_eccflow_options o;
    ssk_eccflow_options_default(&o);
    o.input_smooth_sigma = opts_.input_smooth_sigma; o.reference_smooth_sigma = opts_.reference_smooth_sigma;
    o.update_multiplier = opts_.update_multiplier; o.scale_factor = opts_.scale_factor; o.noise_level = opts_.noise_level;
    o.max_iterations = opts_.max_iterations; o.support_scale = opts_.support_scale; o.min_image_size = opts_.min_image_size;
    o.max_pyramid_level = opts_.max_pyramid_level; o.downscale_method = opts_.downscale;
    if (ssk_eccflow_create(&o, &h_) != SSK_OK) return false;
    ssk_mat im = detail::view(reference_image);
    detail::Opt<image_t> mk(reference_mask);
    rows_ = im.rows; cols_ = im.cols;
    return ssk_eccflow_set_reference_image(h_, &im, mk.get()) == SSK_OK;
  }
  // compute(input_image, rmap, input_mask): an empty rmap starts from a zero flow (ecc2.cc:2783-2786)
  bool compute(const image_t &input_image, image_t &rmap, const image_t &input_mask = image_t()) {
    if (!h_) return false;
    const int have = rmap.empty() ? 0 : 1;
    if (!have) create_like(rmap, rows_, cols_, SSK_32FC2);
    ssk_mat im = detail::view(input_image), rm = detail::view(rmap);
    detail::Opt<image_t> mk(input_mask);
    return ssk_eccflow_compute(h_, &im, mk.get(), &rm, have) == SSK_OK;
  }
  // the reference tells this overload from the one above by the map's type (cv::Mat2f &); with one image type the input mask
  // is spelled out here (an empty image = cv::noArray())
  bool compute(const image_t &input_image, const image_t &reference_image, image_t &rmap, const image_t &input_mask,
               const image_t &reference_mask = image_t()) {
    return set_reference_image(reference_image, reference_mask) && compute(input_image, rmap, input_mask);
  }
  bool current_uv(image_t &uv) const {
    if (!h_) return false;
    create_like(uv, rows_, cols_, SSK_32FC2);
    ssk_mat v = detail::view(uv);
    return ssk_eccflow_get_uv(h_, &v) == SSK_OK;
  }
  int num_levels() const { return h_ ? ssk_eccflow_num_levels(h_) : 0; }

 private:
  ssk_eccflow *h_ = nullptr;
  c_eccflow_options opts_;
  int rows_ = 0, cols_ = 0;
};

// ---------------------------------------------------------------------------------------------------------
// c_frame_registration, ECC branch  (core/proc/image_registration/c_frame_registration.h:210-354)
// ---------------------------------------------------------------------------------------------------------
// c_ecc_registration_options / c_image_registration_options (c_frame_registration.h:47-64, 119-136): the reference's field
// names and defaults, so that option plumbing written against the reference compiles unchanged.  The sparse-feature stage is
// not part of this library, and the eccflow stage runs after the ECC stage only: setup_reference_frame() fails when
// enable_ecc_registration is off.
struct c_ecc_registration_options {
  double scale = 0.5;
  double eps = 0.2;
  double min_rho = 0.8;
  double input_smooth_sigma = 1.0;
  double reference_smooth_sigma = 1.0;
  double update_step_scale = 1.5;
  int se_radius = 5;
  int ecc_method = SSK_ECC_LM;
  int max_iterations = 50;
  int ecch_max_level = 0;
  int ecch_minimum_image_size = 16;
  double normalization_noise = 0.01;
  int normalization_scale = 0;
  bool ecch_estimate_translation_first = true;
  bool replace_planetary_disk_with_mask = false;
};

// c_eccflow_registration_options (c_frame_registration.h:88-100)
struct c_eccflow_registration_options {
  double update_multiplier = 1.5;
  double input_smooth_sigma = 0;
  double reference_smooth_sigma = 0;
  double noise_level = -1;
  double scale_factor = 0.75;
  int max_iterations = 3;
  int support_scale = 4;
  int min_image_size = -1;
  int max_pyramid_level = -1;
  int downscale_method = SSK_ECCFLOW_DOWNSCALE_RECURSIVE_RESIZE;
};

struct c_image_registration_options {
  int motion_type = SSK_MOTION_AFFINE;
  int ecc_registration_channel = 0;          // color_channel_gray
  int interpolation = SSK_INTER_LINEAR;
  int border_mode = SSK_BORDER_REFLECT101;
  double border_value[4] = {0, 0, 0, 0};
  bool enable_feature_registration = true;
  bool enable_ecc_registration = false;
  bool enable_eccflow_registration = false;
  bool accumulate_and_compensate_turbulent_flow = false;
  c_ecc_registration_options ecc;
  c_eccflow_registration_options eccflow;
};

inline ssk_registration_options to_ssk_options(const c_image_registration_options &o) {
  ssk_registration_options r;
  ssk_registration_options_default(&r);
  r.motion_type = o.motion_type; r.interpolation = o.interpolation; r.border_mode = o.border_mode;
  for (int i = 0; i < 4; ++i) r.border_value[i] = o.border_value[i];
  r.enable_ecc_registration = o.enable_ecc_registration ? 1 : 0;
  r.ecc.scale = o.ecc.scale; r.ecc.eps = o.ecc.eps; r.ecc.min_rho = o.ecc.min_rho;
  r.ecc.input_smooth_sigma = o.ecc.input_smooth_sigma; r.ecc.reference_smooth_sigma = o.ecc.reference_smooth_sigma;
  r.ecc.update_step_scale = o.ecc.update_step_scale; r.ecc.se_radius = o.ecc.se_radius; r.ecc.ecc_method = o.ecc.ecc_method;
  r.ecc.max_iterations = o.ecc.max_iterations; r.ecc.ecch_max_level = o.ecc.ecch_max_level;
  r.ecc.ecch_minimum_image_size = o.ecc.ecch_minimum_image_size; r.ecc.normalization_noise = o.ecc.normalization_noise;
  r.ecc.normalization_scale = o.ecc.normalization_scale;
  r.ecc.ecch_estimate_translation_first = o.ecc.ecch_estimate_translation_first ? 1 : 0;
  r.ecc.replace_planetary_disk_with_mask = o.ecc.replace_planetary_disk_with_mask ? 1 : 0;
  // c_frame_registration.cc:637-660
  r.enable_eccflow_registration = o.enable_eccflow_registration ? 1 : 0;
  r.eccflow.update_multiplier = o.eccflow.update_multiplier; r.eccflow.input_smooth_sigma = o.eccflow.input_smooth_sigma;
  r.eccflow.reference_smooth_sigma = o.eccflow.reference_smooth_sigma; r.eccflow.noise_level = o.eccflow.noise_level;
  r.eccflow.scale_factor = o.eccflow.scale_factor; r.eccflow.max_iterations = o.eccflow.max_iterations;
  r.eccflow.support_scale = o.eccflow.support_scale; r.eccflow.min_image_size = o.eccflow.min_image_size;
  r.eccflow.max_pyramid_level = o.eccflow.max_pyramid_level; r.eccflow.downscale_method = o.eccflow.downscale_method;
  return r;
}

struct c_image_registration_status { ssk_ecc_status ecc = {}; };

class c_frame_registration {
 public:
  typedef std::shared_ptr<c_frame_registration> sptr;
  c_frame_registration() { ssk_registration_options_default(&opts_); }
  explicit c_frame_registration(const ssk_registration_options &o) : opts_(o) {}
  // the reference's constructor (c_frame_registration.h:216): c_frame_registration(const c_image_registration_options &)
  explicit c_frame_registration(const c_image_registration_options &o)
      : opts_(to_ssk_options(o)), other_stage_only_((o.enable_feature_registration || o.enable_eccflow_registration) && !o.enable_ecc_registration) {}
  ~c_frame_registration() { if (h_) ssk_reg_destroy(h_); }
  c_frame_registration(const c_frame_registration &) = delete;
  c_frame_registration &operator=(const c_frame_registration &) = delete;

  ssk_registration_options &options() { return opts_; }
  const c_image_transform::sptr &image_transform() const { return transform_; }
  const c_image_registration_status &status() const { return status_; }
  void set_input_bpp(int bpp) { bpp_ = bpp; }   // 8U/16U frames: c_image_stacking_pipeline_base.cc:271-276 scaling
  // frame timestamps (c_frame_registration.h:230-236): carried for the caller, the ECC stages do not read them
  void set_current_timestamp(double v, bool valid) { current_ts_ = v; current_ts_valid_ = valid; }
  double current_timestamp() const { return current_ts_; }
  bool has_valid_current_timestamp() const { return current_ts_valid_; }
  void set_reference_timestamp(double v, bool valid) { reference_ts_ = v; reference_ts_valid_ = valid; }
  double reference_timestamp() const { return reference_ts_; }
  bool has_valid_reference_timestamp() const { return reference_ts_valid_; }

  bool setup_reference_frame(const image_t &image, const image_t &msk = image_t()) {
    if (other_stage_only_) return false;   // sparse-feature / eccflow registration without the ECC stage: not in this library
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
  bool other_stage_only_ = false;
  double current_ts_ = 0, reference_ts_ = 0;
  bool current_ts_valid_ = false, reference_ts_valid_ = false;
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
  // compute(avg, mask, dscale, ddepth) (c_frame_accumulation.h:22): ddepth < 0 or CV_32F keeps the accumulator's float
  // samples; CV_8U / CV_16U round and saturate like cv::Mat::convertTo
  bool compute(image_t &avg, image_t *mask = nullptr, double dscale = 1.0, int ddepth = -1) const {
    int cols = 0, rows = 0, cn = 0;
    if (ssk_acc_size(h_, &cols, &rows, &cn) != SSK_OK || cols <= 0) return false;
    if (ddepth >= 0 && ddepth != SSK_32F && ddepth != SSK_8U && ddepth != SSK_16U) return false;
    image_t favg;
    image_t &out = (ddepth < 0 || ddepth == SSK_32F) ? avg : favg;
    create_like(out, rows, cols, SSK_MAKETYPE(SSK_32F, cn));
    ssk_mat a = detail::view(out), m;
    if (mask) { create_like(*mask, rows, cols, SSK_8UC1); m = detail::view(*mask); }
    if (ssk_acc_compute(h_, &a, mask ? &m : nullptr, dscale) != SSK_OK) return false;
    if (&out == &favg) {
      create_like(avg, rows, cols, SSK_MAKETYPE(ddepth, cn));
      const ssk_mat s = detail::view(favg), d = detail::view(avg);
      const double hi = ddepth == SSK_8U ? 255.0 : 65535.0;
      for (int y = 0; y < rows; ++y) {
        const float *sp = reinterpret_cast<const float *>(static_cast<const char *>(s.data) + y * s.step);
        for (int x = 0; x < cols * cn; ++x) {
          const double v = std::nearbyint((double)sp[x]);
          const double c = v < 0 ? 0 : v > hi ? hi : v;
          if (ddepth == SSK_8U) reinterpret_cast<uint8_t *>(static_cast<char *>(d.data) + y * d.step)[x] = (uint8_t)c;
          else reinterpret_cast<uint16_t *>(static_cast<char *>(d.data) + y * d.step)[x] = (uint16_t)c;
        }
      }
    }
    return true;
  }
  // get_acc_counters(accw) (c_frame_accumulation.h:24): the weight sums (Bayer: per-colour counters, G halved)
  bool get_acc_counters(image_t &accw) const {
    int cols = 0, rows = 0, cn = 0;
    if (ssk_acc_size(h_, &cols, &rows, &cn) != SSK_OK || cols <= 0) return false;
    create_like(accw, rows, cols, SSK_MAKETYPE(SSK_32F, counter_channels()));
    ssk_mat v = detail::view(accw);
    return ssk_acc_get_counters(h_, &v) == SSK_OK;
  }
  // Multi-GPU: combine the accumulators of all ranks on `root` (ssk_acc_reduce; nccl_comm is the caller's ncclComm_t)
  bool reduce(void *nccl_comm, int root = 0) { return ssk_acc_reduce(h_, nccl_comm, root) == SSK_OK; }
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
  virtual int counter_channels() const { return 1; }
  ssk_acc *h_ = nullptr;
};

class c_weigthed_average : public c_frame_accumulation {   // (sic) reference spelling
 public:
  c_weigthed_average() : c_frame_accumulation(SSK_ACC_WEIGHTED_AVERAGE) {}
  // accumulator() / counter() (c_frame_accumulation.h:58-59): copies of the running mean and of the weight sums
  bool accumulator(image_t &acc) const { return compute(acc, nullptr, 1.0, -1); }
  bool counter(image_t &cntr) const { return get_acc_counters(cntr); }
};

class c_bayer_average : public c_frame_accumulation {
 public:
  c_bayer_average() : c_frame_accumulation(SSK_ACC_BAYER_AVERAGE) {}
  void set_bayer_pattern(int colorid) { ssk_acc_set_bayer_pattern(h_, colorid); }
  bool set_remap(const image_t &rmap) {
    rmap_ = rmap;
    if (rmap.empty()) return ssk_acc_set_remap(h_, nullptr, nullptr) == SSK_OK;
    ssk_mat v = detail::view(rmap);
    return ssk_acc_set_remap(h_, nullptr, &v) == SSK_OK;
  }
  const image_t &remap() const { return rmap_; }     // c_bayer_average::remap() (c_frame_accumulation.h:249)

 protected:
  int counter_channels() const override { return 3; }

 private:
  image_t rmap_;
};

// c_local_variance_sharpness_measure::compute (c_local_variance_sharpness_measure.cc:193-247)
// c_canvas_average (core/average/c_frame_accumulation.h:65-137): boxes are {x, y, width, height}
class c_canvas_average {
 public:
  struct options { int interpolation = SSK_INTER_LINEAR; } opts;
  c_canvas_average() = default;
  ~c_canvas_average() { if (h_) ssk_canvas_destroy(h_); }
  c_canvas_average(const c_canvas_average &) = delete;
  c_canvas_average &operator=(const c_canvas_average &) = delete;
  void setCanvasSize(int cols, int rows) { want_cols_ = cols; want_rows_ = rows; clear(); }
  int accumulated_frames() const { return h_ ? ssk_canvas_accumulated_frames(h_) : 0; }
  void accumulator_size(int *cols, int *rows) const { *cols = *rows = 0; if (h_) ssk_canvas_size(h_, cols, rows, nullptr); }
  void last_bbox(int bbox[4]) const { bbox[0] = bbox[1] = bbox[2] = bbox[3] = 0; if (h_) ssk_canvas_last_bbox(h_, bbox); }
  bool add(const image_t &current_image, const image_t &current_weights_or_mask = image_t(), const image_t &rmap = image_t(),
           const int *new_canvas_bbox = nullptr) {
    if (!h_) {   // opts.interpolation is latched with the first frame
      if (ssk_canvas_create(opts.interpolation, &h_) != SSK_OK) return false;
      ssk_canvas_set_canvas_size(h_, want_cols_, want_rows_);
    }
    ssk_mat im = detail::view(current_image);
    detail::Opt<image_t> w(current_weights_or_mask), m(rmap);
    return ssk_canvas_add(h_, &im, w.get(), m.get(), new_canvas_bbox) == SSK_OK;
  }
  bool compute(image_t &avg, image_t *mask = nullptr, double dscale = 1.0, int ddepth = -1, const int *rbbox = nullptr) const {
    if (!h_ || (ddepth >= 0 && ddepth != SSK_32F)) return false;
    int cols = 0, rows = 0, cn = 0;
    ssk_canvas_size(h_, &cols, &rows, &cn);
    int x0 = 0, y0 = 0, x1 = cols, y1 = rows;
    if (rbbox && rbbox[2] > 0 && rbbox[3] > 0) {
      x0 = rbbox[0] > 0 ? rbbox[0] : 0; y0 = rbbox[1] > 0 ? rbbox[1] : 0;
      x1 = rbbox[0] + rbbox[2] < cols ? rbbox[0] + rbbox[2] : cols; y1 = rbbox[1] + rbbox[3] < rows ? rbbox[1] + rbbox[3] : rows;
    }
    if (x1 <= x0 || y1 <= y0) return false;
    create_like(avg, y1 - y0, x1 - x0, SSK_MAKETYPE(SSK_32F, cn));
    ssk_mat a = detail::view(avg), mk;
    if (mask) { create_like(*mask, y1 - y0, x1 - x0, SSK_8UC1); mk = detail::view(*mask); }
    return ssk_canvas_compute(h_, &a, mask ? &mk : nullptr, dscale, rbbox) == SSK_OK;
  }
  void clear() { if (h_) { ssk_canvas_destroy(h_); h_ = nullptr; } }
  static void computeCanvasSize(int frame_cols, int frame_rows, int *cols, int *rows) { *cols = 3 * frame_cols / 2; *rows = 3 * frame_rows / 2; }

 private:
  ssk_canvas *h_ = nullptr;
  int want_cols_ = 0, want_rows_ = 0;
};

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

// c_jovian_derotation_remap / c_saturn_derotation_remap (core/proc/feature2d/c_jovian_derotation_remap.{h,cc},
// c_saturn_derotation_remap.{h,cc}): the two classes are the same code in the reference apart from the default rotation
// period; pose bookkeeping on the host (ssk_build_ellipsoid_rotation / ssk_ellipsoid_bbox), the map on the device.
struct c_lpg_options { double k = 2.0, p = 2.0; int dscale = 2, uscale = 6; };   // core/proc/lpg.h

template <int PERIOD_MS>
class c_ellipsoid_derotation_remap {
 public:
  static constexpr double default_rotation_period_sec = PERIOD_MS * 1e-3;
  void set_rotation_period_sec(double v) { period_ = v; }
  double rotation_period_sec() const { return period_; }
  bool set_reference_pose(int image_cols, int image_rows, const double center[2], const double axes[3], const double pose[3]) {
    cols_ = image_cols; rows_ = image_rows;
    for (int i = 0; i < 2; ++i) center_[i] = center[i];
    for (int i = 0; i < 3; ++i) { axes_[i] = axes[i]; current_pose_[i] = target_pose_[i] = pose[i]; }
    if (ssk_build_ellipsoid_rotation(target_pose_, Rtarget_) != SSK_OK) return false;
    for (int i = 0; i < 9; ++i) Rcurrent_[i] = Rtarget_[i];
    return ssk_ellipsoid_bbox(rows_, cols_, center_, axes_, Rtarget_, ebox_, crop_box_) == SSK_OK;
  }
  bool compute_derotation_for_angle(double longitude_rotation_radians, double wscale = 1) {
    set_current(longitude_rotation_radians);
    return compute_ellipsoid_zrotation_remap(rows_, cols_, center_, axes_, Rcurrent_, Rtarget_, ebox_[4], crop_box_, wscale, rmap_, wmap_, rmask_);
  }
  bool compute_derotation_for_time(double deltat_sec, double wscale = 1) { return compute_derotation_for_angle(angle_for_time(deltat_sec), wscale); }
  // one frame of derotate_and_average_frames after preproc_align_and_remap (c_jdr_pipeline.cc:1184-1236,
  // c_sdr_pipeline.cc:1192-1246): compute_derotation_for_time(deltat_sec, wscale) and every statement up to
  // _frame_average.add(current_frame, current_weights) as one device chain
  bool derotate_and_add(c_frame_accumulation &acc, const image_t &frame, const image_t &mask, double deltat_sec, double wscale, bool is_master,
                        bool enable_weighted_average = true, const c_lpg_options &lpg = c_lpg_options()) {
    set_current(angle_for_time(deltat_sec));
    ssk_mat f = detail::view(frame);
    detail::Opt<image_t> m(mask);
    return ssk_jdr_derotate_and_add(acc.handle(), &f, m.get(), center_, axes_, Rcurrent_, Rtarget_, ebox_[4], crop_box_, wscale, is_master,
                                    enable_weighted_average, lpg.k, lpg.p, lpg.dscale, lpg.uscale) == SSK_OK;
  }
  const image_t &rmap() const { return rmap_; }
  const image_t &wmap() const { return wmap_; }
  const image_t &rmask() const { return rmask_; }
  const double *center() const { return center_; }
  const double *axes() const { return axes_; }
  const double *current_pose() const { return current_pose_; }
  const double *target_pose() const { return target_pose_; }
  const double *Rcurrent() const { return Rcurrent_; }
  const double *Rtarget() const { return Rtarget_; }
  const float *ebox() const { return ebox_; }          // {center.x, center.y, width, height, angle_deg}
  const int *crop_box() const { return crop_box_; }

 private:
  double angle_for_time(double deltat_sec) const {
    const double period = period_ > 0 ? period_ : default_rotation_period_sec;
    return 2 * 3.1415926535897932384626433832795 * deltat_sec / period;
  }
  void set_current(double angle) {
    current_pose_[0] = target_pose_[0] + angle; current_pose_[1] = target_pose_[1]; current_pose_[2] = target_pose_[2];
    ssk_build_ellipsoid_rotation(current_pose_, Rcurrent_);
  }
  double period_ = default_rotation_period_sec;
  int rows_ = 0, cols_ = 0, crop_box_[4] = {0, 0, 0, 0};
  double center_[2] = {0, 0}, axes_[3] = {1, 1, 1}, current_pose_[3] = {0, 0, 0}, target_pose_[3] = {0, 0, 0}, Rcurrent_[9] = {}, Rtarget_[9] = {};
  float ebox_[5] = {};
  image_t rmap_, wmap_, rmask_;
};
using c_jovian_derotation_remap = c_ellipsoid_derotation_remap<35740632>;   // 9h 55m 40.632s (c_jovian_derotation_remap.cc:39)
using c_saturn_derotation_remap = c_ellipsoid_derotation_remap<38018000>;   // 10h 33m 38s (c_saturn_derotation_remap.cc:34)

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
  bool reset() { return ssk_stack_reset(h_) == SSK_OK; }
  // end of a run sharded over ranks: the accumulators of all ranks combined on `root` (ncclComm_t of the caller)
  bool reduce(void *nccl_comm, int root = 0) { return ssk_stack_reduce(h_, nccl_comm, root) == SSK_OK; }
  bool flush() { return ssk_stack_flush(h_) == SSK_OK; }   // stream-side join of the side-stream ring kernel (see ssk.h)

 private:
  ssk_stack *h_ = nullptr;
};


// linear_interpolation_inpaint (core/proc/inpaint/linear_interpolation_inpaint.cc:327-368)
inline bool linear_interpolation_inpaint(const image_t &src, const image_t &mask, image_t &dst) {
  ssk_mat s = detail::view(src);
  image_t out;
  create_like(out, s.rows, s.cols, s.type);
  ssk_mat d = detail::view(out);
  detail::Opt<image_t> m(mask);
  const bool ok = ssk_linear_interpolation_inpaint(&s, m.get(), &d) == SSK_OK;
  if (ok) dst = out;
  return ok;
}

// Stream-ordered call chains for device-resident matrices (include/ssk.h): lpg / GaussianBlur / c_frame_accumulation::add /
// derotate_and_add return once enqueued and are ordered on the device by the library; compute() waits for the chain.
inline bool set_stream_ordered(bool enable) { return ssk_set_stream_ordered(enable ? 1 : 0) != 0; }   // returns the previous mode
inline bool device_synchronize() { return ssk_device_synchronize() == SSK_OK; }

// median_filter_bad_pixels (core/proc/bad_pixels.cc:58-70): in place; a Bayer colorid dispatches to bayer_denoise like the reference
inline bool bayer_denoise(image_t &image, double variation_threshold) {   // core/io/debayer.cc:1599-1611, returnBayerPlanes = false
  ssk_mat m = detail::view(image);
  return ssk_bayer_denoise(&m, variation_threshold) == SSK_OK;
}
inline bool median_filter_bad_pixels(image_t &image, double variation_threshold, bool is_bayer_pattern = false) {
  if (is_bayer_pattern) return bayer_denoise(image, variation_threshold);
  ssk_mat m = detail::view(image);
  return ssk_median_filter_bad_pixels(&m, variation_threshold) == SSK_OK;
}

// average_bayer_planes (core/io/debayer.cc:277-376), raw single-channel form
inline bool average_bayer_planes(const image_t &src, image_t &dst) {
  ssk_mat s = detail::view(src);
  image_t out;
  create_like(out, s.rows / 2, s.cols / 2, s.type);
  ssk_mat d = detail::view(out);
  const bool ok = ssk_average_bayer_planes(&s, &d) == SSK_OK;
  if (ok) dst = out;
  return ok;
}

// c_ser_reader (core/io/c_ser_file.h:136-190)
class c_ser_reader {
 public:
  c_ser_reader() = default;
  explicit c_ser_reader(const std::string &filename) { open(filename); }
  ~c_ser_reader() { close(); }
  c_ser_reader(const c_ser_reader &) = delete;
  c_ser_reader &operator=(const c_ser_reader &) = delete;
  bool open(const std::string &filename) {
    close();
    if (ssk_ser_open(filename.c_str(), &h_) != SSK_OK) return false;
    return ssk_ser_info(h_, &cols_, &rows_, &type_, &bpp_, &color_id_, &frames_, &has_ts_) == SSK_OK;
  }
  void close() { if (h_) { ssk_ser_close(h_); h_ = nullptr; } curpos_ = 0; }
  bool is_open() const { return h_ != nullptr; }
  int image_width() const { return cols_; }
  int image_height() const { return rows_; }
  int bits_per_plane() const { return bpp_; }
  int color_id() const { return color_id_; }
  int num_frames() const { return frames_; }
  int curpos() const { return curpos_; }
  bool seek(int frame_index) { if (frame_index < 0) frame_index = 0; if (frame_index >= frames_) return false; curpos_ = frame_index; return true; }
  bool read(image_t &image, uint64_t *timestamp = nullptr) {
    if (!h_ || curpos_ >= frames_) return false;
    create_like(image, rows_, cols_, type_);
    ssk_mat v = detail::view(image);
    if (ssk_ser_read(h_, curpos_, &v, timestamp) != SSK_OK) return false;
    ++curpos_;
    return true;
  }

 private:
  ssk_ser *h_ = nullptr;
  int cols_ = 0, rows_ = 0, type_ = 0, bpp_ = 0, color_id_ = 0, frames_ = 0, has_ts_ = 0, curpos_ = 0;
};

// ---------------------------------------------------------------------------------------------------------
// c_image_stacking_pipeline, the stacking part of run_pipeline (c_image_stacking_pipeline.cc:436-466, 731-769, 1112-1312,
// 1358-1862) over an in-memory sequence: master frame (selected frame alone, or generated from the frames around it),
// unsharp_mask of the master, the batched per-frame loop, compute() + average_pyramid_inpaint.  Option names follow
// c_image_stacking_master_options / c_frame_accumulation_options / c_image_stacking_options (c_image_stacking_pipeline.h:88-182).
// ---------------------------------------------------------------------------------------------------------
struct c_frame_accumulation_options {
  int accumulation_method = SSK_STACK_AVERAGE;
  struct { int dscale = 1, kradius = 1, uscale = 0; } sharpness_measure;
};

// c_frame_upscale_options (c_image_stacking_pipeline.h:33-86)
enum frame_upscale_stage { frame_upscale_stage_unknown = -1, frame_upscale_after_align = 1, frame_upscale_before_align = 2 };
enum frame_upscale_option { frame_upscale_none = 0, frame_upscale_pyrUp = 1, frame_upscale_x15 = 2, frame_upscale_x30 = 3 };
struct c_frame_upscale_options {
  frame_upscale_option upscale_option = frame_upscale_none;
  frame_upscale_stage upscale_stage = frame_upscale_after_align;
  bool need_upscale_before_align() const { return upscale_option != frame_upscale_none && upscale_stage == frame_upscale_before_align; }
  bool need_upscale_after_align() const { return upscale_option != frame_upscale_none && upscale_stage == frame_upscale_after_align; }
  double image_scale() const { return upscale_option == frame_upscale_x15 ? 1.5 : upscale_option == frame_upscale_pyrUp ? 2 : upscale_option == frame_upscale_x30 ? 3 : 1.0; }
};

// c_image_stacking_pipeline::upscale_image / upscale_remap / upscale_optflow (c_image_stacking_pipeline.cc:1869-2002)
inline bool upscale_image(frame_upscale_option scale, const image_t &src, const image_t &srcmask, image_t &dst, image_t *dstmask = nullptr) {
  ssk_mat s = detail::view(src);
  int uc = 0, ur = 0;
  if (ssk_upscale_size(scale, s.cols, s.rows, &uc, &ur) != SSK_OK) return false;
  image_t out, outm;                      // the reference up-scales in place: src and dst may be the same image
  create_like(out, ur, uc, s.type);
  ssk_mat d = detail::view(out), sm, dm;
  const bool with_mask = dstmask && !srcmask.empty();
  if (with_mask) { create_like(outm, ur, uc, SSK_8UC1); sm = detail::view(srcmask); dm = detail::view(outm); }
  if (ssk_upscale_image(scale, &s, with_mask ? &sm : nullptr, &d, with_mask ? &dm : nullptr) != SSK_OK) return false;
  dst = out;
  if (with_mask) *dstmask = outm;
  return true;
}
inline bool upscale_remap(frame_upscale_option scale, const image_t &srcmap, image_t &dstmap) {
  ssk_mat s = detail::view(srcmap);
  int uc = 0, ur = 0;
  if (ssk_upscale_size(scale, s.cols, s.rows, &uc, &ur) != SSK_OK) return false;
  image_t out;
  create_like(out, ur, uc, SSK_32FC2);
  ssk_mat d = detail::view(out);
  if (ssk_upscale_remap(scale, &s, &d) != SSK_OK) return false;
  dstmap = out;
  return true;
}
inline bool upscale_optflow(frame_upscale_option scale, const image_t &srcmap, image_t &dstmap) {
  ssk_mat s = detail::view(srcmap);
  int uc = 0, ur = 0;
  if (ssk_upscale_size(scale, s.cols, s.rows, &uc, &ur) != SSK_OK) return false;
  image_t out;
  create_like(out, ur, uc, SSK_32FC2);
  ssk_mat d = detail::view(out);
  if (ssk_upscale_optflow(scale, &s, &d) != SSK_OK) return false;
  dstmap = out;
  return true;
}

struct c_image_stacking_master_options {
  c_image_registration_options registration;
  c_frame_accumulation_options accumulation;
  int master_frame_index = 0;
  int max_frames_to_generate_master_frame = 3000;
  bool generate_master_frame = true;
  double unsharp_sigma = 1.0, unsharp_alpha = 0.8;
};

class c_image_stacking_pipeline {
 public:
  c_image_stacking_master_options master_options;
  c_image_registration_options registration_options;   // of the stacking pass
  c_frame_accumulation_options accumulation_options;
  c_frame_upscale_options upscale_options;              // frame_upscale_after_align is fused into the loop; before_align up-scales
                                                        // the frames first (weights are then computed on the up-scaled frames: the
                                                        // reference computes them before up-scaling, c_image_stacking_pipeline.cc:1465-1496)
  int bayer_colorid = SSK_COLORID_BAYER_RGGB;
  int max_batch = 32;

  // frames: the whole input sequence (same size and type); bpp: bits per sample of integer frames
  bool run(const std::vector<image_t> &frames, int bpp, image_t &stacked, image_t &stacked_mask, bool inpaint = true) {
    if (frames.empty()) return false;
    image_t reference, refmask;
    if (!create_reference_frame(frames, bpp, reference, refmask)) return false;
    ssk_stack_options so = stack_options(registration_options, accumulation_options, false);
    if (upscale_options.need_upscale_after_align()) { so.upscale_option = upscale_options.upscale_option; so.upscale_stage = SSK_UPSCALE_AFTER_ALIGN; }
    if (upscale_options.need_upscale_before_align()) return false;   // see upscale_options above: up-scale the frames with upscale_image() first
    c_stacking_loop loop(so);
    if (!loop.valid() || !loop.set_reference(reference, bpp)) return false;
    if (!add_all(loop, frames, 0, (int)frames.size(), bpp)) return false;
    accumulated_frames_ = loop.accumulated_frames();
    const ssk_mat r = detail::view(reference);
    const int cn = accumulation_options.accumulation_method == SSK_STACK_BAYER_AVERAGE ? 3 : ((r.type >> 3) + 1);
    int orows = r.rows, ocols = r.cols;
    if (so.upscale_option != SSK_UPSCALE_NONE && so.enable_registration) ssk_upscale_size(so.upscale_option, r.cols, r.rows, &ocols, &orows);
    return inpaint ? loop.compute_inpainted(stacked, stacked_mask, orows, ocols, cn) : loop.compute(stacked, stacked_mask, orows, ocols, cn);
  }
  // create_reference_frame (c_image_stacking_pipeline.cc:1112-1312)
  bool create_reference_frame(const std::vector<image_t> &frames, int bpp, image_t &reference, image_t &refmask) {
    const int n = (int)frames.size();
    const int pos = master_options.master_frame_index < 0 ? 0 : master_options.master_frame_index >= n ? n - 1 : master_options.master_frame_index;
    const int max_stack = master_options.generate_master_frame ? master_options.max_frames_to_generate_master_frame : 1;
    const ssk_mat f0 = detail::view(frames[pos]);
    if (max_stack < 2 || n < 2) {
      // the selected frame alone, converted like read_input_frame does (CV_32F, 1 / (1 << bpp))
      create_like(reference, f0.rows, f0.cols, SSK_MAKETYPE(SSK_32F, (f0.type >> 3) + 1));
      ssk_mat d = detail::view(reference);
      if (ssk_input_calibrate(&f0, bpp, nullptr, nullptr, &d) != SSK_OK) return false;
    } else {
      ssk_stack_options so = stack_options(master_options.registration, master_options.accumulation, true);
      c_stacking_loop loop(so);
      if (!loop.valid() || !loop.set_reference(frames[pos], bpp)) return false;
      int lo, hi;
      master_frame_range(n, pos, max_stack, &lo, &hi);
      if (!add_all(loop, frames, lo, hi, bpp) || loop.accumulated_frames() < 1) return false;
      const int cn = master_options.accumulation.accumulation_method == SSK_STACK_BAYER_AVERAGE ? 3 : ((f0.type >> 3) + 1);
      image_t avg;
      if (!loop.compute(avg, refmask, f0.rows, f0.cols, cn)) return false;
      if (!linear_interpolation_inpaint(avg, refmask, reference)) return false;
    }
    if (master_options.unsharp_sigma > 0 && master_options.unsharp_alpha > 0)
      return unsharp_mask(reference, reference, master_options.unsharp_sigma, master_options.unsharp_alpha);
    return true;
  }
  // [startpos, endpos) of the frames stacked into the master frame (c_image_stacking_pipeline.cc:1211-1229)
  static void master_frame_range(int num_frames, int master_frame_pos, int max_frames_to_stack, int *startpos, int *endpos) {
    if (max_frames_to_stack >= num_frames) { *startpos = 0; *endpos = num_frames; return; }
    int s = master_frame_pos - max_frames_to_stack / 2;
    if (s < 0) s = 0;
    int e = s + max_frames_to_stack;
    if (e >= num_frames) { s = num_frames - max_frames_to_stack; if (s < 0) s = 0; e = num_frames; }
    *startpos = s; *endpos = e;
  }
  int accumulated_frames() const { return accumulated_frames_; }

 private:
  ssk_stack_options stack_options(const c_image_registration_options &r, const c_frame_accumulation_options &a, bool master) const {
    ssk_stack_options so;
    ssk_stack_options_default(&so);
    so.registration = to_ssk_options(r);
    so.enable_registration = r.enable_ecc_registration ? 1 : 0;
    so.registration.enable_ecc_registration = 1;
    so.accumulation_method = a.accumulation_method;
    so.sm_dscale = a.sharpness_measure.dscale; so.sm_kradius = a.sharpness_measure.kradius; so.sm_uscale = a.sharpness_measure.uscale;
    so.bayer_colorid = bayer_colorid;
    so.max_batch = max_batch;
    so.generating_master_frame = master ? 1 : 0;
    return so;
  }
  bool add_all(c_stacking_loop &loop, const std::vector<image_t> &frames, int lo, int hi, int bpp) {
    for (int i = lo; i < hi; i += max_batch) {
      std::vector<image_t> chunk(frames.begin() + i, frames.begin() + (i + max_batch < hi ? i + max_batch : hi));
      if (!loop.add_frames(chunk, bpp)) return false;
    }
    return true;
  }
  int accumulated_frames_ = 0;
};

}  // namespace ssk
