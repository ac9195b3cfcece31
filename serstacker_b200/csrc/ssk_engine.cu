// Engine classes: device-resident c_ecch / c_frame_registration / accumulator state and batching.
#include "ssk_engine.cuh"
#include "ssk_eccflow.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ssk {

int DevBuf::ensure(size_t n) {
  if (n <= bytes && p) return SSK_OK;
  release();
  SSK_CUDA(cudaMalloc(&p, n ? n : 1));
  bytes = n;
  return SSK_OK;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}
int PinnedBuf::ensure(size_t n) {
  if (n <= bytes && p) return SSK_OK;
  if (p) cudaFreeHost(p);
  p = nullptr;
  SSK_CUDA(cudaMallocHost(&p, n ? n : 1));
  bytes = n;
  return SSK_OK;
}

int make_transform(ssk_transform *t, int motion_type) {
  // image_transform.cc:34-63 followed by reset()
  memset(t, 0, sizeof(*t));
  t->motion_type = motion_type;
  t->aux[2] = 1.f;   // homography a22
  t->aux[3] = 1.f;   // euclidean scale when fixed
  switch (motion_type) {
    case SSK_MOTION_TRANSLATION: t->nparams = 2; break;
    case SSK_MOTION_EUCLIDEAN: t->nparams = 3; break;
    case SSK_MOTION_SCALED_EUCLIDEAN: t->nparams = 4; t->params[3] = 1.f; break;
    case SSK_MOTION_AFFINE: t->nparams = 6; t->params[0] = 1.f; t->params[4] = 1.f; break;
    case SSK_MOTION_HOMOGRAPHY: t->nparams = 8; t->params[0] = 1.f; t->params[4] = 1.f; break;
    default: set_error("unsupported motion type"); return SSK_ERR_INVALID;
  }
  return SSK_OK;
}

int host_scale_transform(ssk_transform *t, double f) {
  // c_image_transform::scale_transfrom (c_image_transform.cc:130-134, 502-507, 911-915, 1186-1194)
  switch (t->motion_type) {
    case SSK_MOTION_TRANSLATION:
      t->params[0] = (float)(t->params[0] * f); t->params[1] = (float)(t->params[1] * f); break;
    case SSK_MOTION_EUCLIDEAN:
    case SSK_MOTION_SCALED_EUCLIDEAN:
      t->params[0] = (float)(t->params[0] * f); t->params[1] = (float)(t->params[1] * f);
      t->aux[0] = (float)(t->aux[0] * f); t->aux[1] = (float)(t->aux[1] * f); break;
    case SSK_MOTION_AFFINE:
      t->params[2] = (float)(t->params[2] * f); t->params[5] = (float)(t->params[5] * f); break;
    case SSK_MOTION_HOMOGRAPHY:
      t->params[2] = (float)(t->params[2] * f); t->params[5] = (float)(t->params[5] * f);
      t->params[6] = (float)(t->params[6] / f); t->params[7] = (float)(t->params[7] / f); break;
    default: set_error("unsupported motion type"); return SSK_ERR_INVALID;
  }
  return SSK_OK;
}

// cv::getGaussianKernel(ksize, sigma) -> float taps (sepFilter2D converts the kernel to the data type)
static int gaussian_kernel(double sigma, float *k) {
  const int n = std::max(3, 2 * ((int)(3 * sigma)) + 1);   // ecc2.cc:1001
  double cf[kMaxTaps], sum = 0;
  const double s2 = -0.5 / (sigma * sigma);
  for (int i = 0; i < n; ++i) {
    const double x = i - (n - 1) * 0.5;
    cf[i] = std::exp(s2 * x * x);
    sum += cf[i];
  }
  for (int i = 0; i < n; ++i) k[i] = (float)(cf[i] / sum);
  return n;
}

static const float kD5[5] = {1.f / 12.f, -2.f / 3.f, 0.f, 2.f / 3.f, -1.f / 12.f};  // ecc2.cc:148
static const float kS3[3] = {0.25f, 0.5f, 0.25f};                                   // ecc2.cc:149

// ------------------------------------------------------------------------------------------------
int mask_to_device(const ssk_mat *mask, int rows, int cols, DevBuf &staging, cudaStream_t s, const uint8_t **d_mask, int64_t *step) {
  SSK_REQUIRE(mask && mask->data, "mask: null data");
  SSK_REQUIRE(mask->type == SSK_8UC1, "mask must be CV_8UC1");
  SSK_REQUIRE(mask->rows == rows && mask->cols == cols, "mask size differs from the image size");
  if (mask->mem == SSK_MEM_DEVICE) {
    *d_mask = static_cast<const uint8_t *>(mask->data); *step = mask->step;
    return SSK_OK;
  }
  if (int e = staging.ensure((size_t)rows * cols)) return e;
  SSK_CUDA(cudaMemcpy2DAsync(staging.p, cols, mask->data, mask->step, cols, rows, cudaMemcpyHostToDevice, s));
  *d_mask = staging.as<uint8_t>(); *step = cols;
  return SSK_OK;
}

Ecch::~Ecch() {}

int Ecch::init(const ssk_ecch_options &o, cudaStream_t s) {
  opts = o;
  stream = s;
  have_reference = false;
  // c_ecc_align::_interpolation (ecc2.h:137) is the warp the forward-additive solver samples the current image with; the
  // other solvers always sample bilinearly.  c_frame_registration never changes the default (INTER_LINEAR)
  SSK_REQUIRE(o.interpolation == SSK_INTER_NEAREST || o.interpolation == SSK_INTER_LINEAR || o.interpolation == SSK_INTER_AREA,
              "c_ecch: interpolation must be NEAREST, LINEAR or AREA (the solvers' image warp)");
  if (const char *e = getenv("SSK_ECC_CLUSTER")) {   // tuning knob: CTAs per frame (thread-block cluster size)
    const int c = atoi(e);
    if (c == 1 || c == 2 || c == 4 || c == 8) { cluster_size = c; cluster_fixed = true; }
  }
  return SSK_OK;
}

int Ecch::set_reference(const float *d_img, int rows, int cols, const uint8_t *d_mask) {
  SSK_REQUIRE(opts.reference_smooth_sigma < 5.0 && opts.input_smooth_sigma < 5.0, "ECC smoothing sigma must be < 5");
  // number of levels: ecc2.cc:1009-1026
  const int min_image_size = std::max(4, opts.minimum_image_size);
  int w = cols, h = rows, lv = 1;
  lw[0] = w; lh[0] = h;
  while (true) {
    if (opts.maxlevel >= 0 && lv >= std::max(0, opts.maxlevel)) break;
    const int nw = ((w + 1) >> 1) & ~1, nh = ((h + 1) >> 1) & ~1;   // ecc2.h:290-293
    if (nw < min_image_size || nh < min_image_size) break;
    if (lv >= kMaxLevels) break;
    w = nw; h = nh;
    lw[lv] = w; lh[lv] = h;
    ++lv;
  }
  nlevels = lv;
  pyr_floats = 0;
  for (int l = 0; l < nlevels; ++l) {
    loff[l] = pyr_floats;
    pyr_floats += ((int64_t)lw[l] * lh[l] + 63) & ~(int64_t)63;   // keep levels 256-byte aligned
  }
  if (int e = ref_pyr.ensure(pyr_floats * 4)) return e;
  float *rp = ref_pyr.as<float>();

  SepFilterArgs sf = {};
  sf.rows = rows; sf.cols = cols; sf.batch = 1;
  sf.src = d_img; sf.dst = rp;
  if (opts.reference_smooth_sigma > 0) {
    gauss_ref_n = gaussian_kernel(opts.reference_smooth_sigma, gauss_ref);
    sf.kxn = sf.kyn = gauss_ref_n;
    memcpy(sf.kx, gauss_ref, sizeof(float) * gauss_ref_n);
    memcpy(sf.ky, gauss_ref, sizeof(float) * gauss_ref_n);
  } else {
    sf.kxn = sf.kyn = 1; sf.kx[0] = sf.ky[0] = 1.f;
  }
  if (int e = launch_sepfilter(sf, stream)) return e;
  for (int l = 1; l < nlevels; ++l) {
    PyrDownArgs pd = {};
    pd.src.data = rp + loff[l - 1]; pd.src.step = (int64_t)lw[l - 1] * 4; pd.src.rows = lh[l - 1]; pd.src.cols = lw[l - 1];
    pd.src.depth = SSK_32F; pd.src.cn = 1; pd.src.scale = 1.f;
    pd.dst = rp + loff[l]; pd.dst_rows = lh[l]; pd.dst_cols = lw[l]; pd.batch = 1; pd.post_scale = 1.f;
    if (int e = launch_pyrdown(pd, stream)) return e;
  }
  const bool ic = opts.method == SSK_ECC_INVERSE_COMPOSITIONAL || opts.method == SSK_ECC_INVERSE_COMPOSITIONAL_LM;
  if (ic) {
    if (int e = ref_gx.ensure(pyr_floats * 4)) return e;
    if (int e = ref_gy.ensure(pyr_floats * 4)) return e;
    for (int l = 0; l < nlevels; ++l) {
      // ecc_differentiate (ecc2.cc:142-169): gx = sepFilter2D(d5 along x, s3 along y), gy = (s3, d5)
      SepFilterArgs g = {};
      g.rows = lh[l]; g.cols = lw[l]; g.batch = 1; g.src = rp + loff[l];
      g.dst = ref_gx.as<float>() + loff[l];
      g.kxn = 5; g.kyn = 3; memcpy(g.kx, kD5, sizeof(kD5)); memcpy(g.ky, kS3, sizeof(kS3));
      if (int e = launch_sepfilter(g, stream)) return e;
      g.dst = ref_gy.as<float>() + loff[l];
      g.kxn = 3; g.kyn = 5; memcpy(g.kx, kS3, sizeof(kS3)); memcpy(g.ky, kD5, sizeof(kD5));
      if (int e = launch_sepfilter(g, stream)) return e;
    }
    if (int e = d_hp_trans.ensure(sizeof(EccHpCache))) return e;
    if (int e = d_hp_main.ensure(sizeof(EccHpCache))) return e;
  }
  // reference mask: pyramid by cv::resize(INTER_NEAREST) (c_ecch::downscale_image, ecc2.cc:972-982); the forward
  // solvers erode it 5x5 (ecc2.cc:1211-1221, 1425-1431), the inverse-compositional ones use it as given and zero the
  // reference gradients under it (ecc2.cc:162-166, 1855-1860)
  have_ref_mask = d_mask != nullptr;
  for (int l = 0; l < nlevels; ++l) rma[l] = (double)lw[l] * lh[l];
  if (have_ref_mask) {
    if (int e = ref_mask.ensure(pyr_floats)) return e;
    if (int e = ref_mask_tmp.ensure((size_t)lw[0] * lh[0])) return e;
    if (int e = d_count.ensure(sizeof(int) * kMaxLevels)) return e;
    uint8_t *mp = ref_mask.as<uint8_t>();
    SSK_CUDA(cudaMemcpyAsync(mp, d_mask, (size_t)lw[0] * lh[0], cudaMemcpyDeviceToDevice, stream));
    for (int l = 1; l < nlevels; ++l)
      if (int e = launch_resize_nearest_u8(mp + loff[l - 1], lh[l - 1], lw[l - 1], mp + loff[l], lh[l], lw[l], stream)) return e;
    SSK_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(int) * kMaxLevels, stream));
    for (int l = 0; l < nlevels; ++l) {
      const int n = lw[l] * lh[l];
      if (!ic) {   // erode in place through the scratch buffer
        SSK_CUDA(cudaMemcpyAsync(ref_mask_tmp.p, mp + loff[l], n, cudaMemcpyDeviceToDevice, stream));
        if (int e = launch_erode5_u8(ref_mask_tmp.as<uint8_t>(), lw[l], mp + loff[l], lw[l], lh[l], lw[l], 1, stream)) return e;
      }
      if (int e = launch_apply_refmask(mp + loff[l], n, ic ? ref_gx.as<float>() + loff[l] : nullptr,
                                       ic ? ref_gy.as<float>() + loff[l] : nullptr, d_count.as<int>() + l, stream)) return e;
    }
    int counts[kMaxLevels];
    SSK_CUDA(cudaMemcpyAsync(counts, d_count.p, sizeof(int) * kMaxLevels, cudaMemcpyDeviceToHost, stream));
    SSK_CUDA(cudaStreamSynchronize(stream));
    for (int l = 0; l < nlevels; ++l) rma[l] = (double)counts[l];
  }
  if (ic) {
    SSK_CUDA(cudaMemsetAsync(d_hp_trans.p, 0, sizeof(EccHpCache), stream));
    SSK_CUDA(cudaMemsetAsync(d_hp_main.p, 0, sizeof(EccHpCache), stream));
  }
  if (opts.input_smooth_sigma > 0) gauss_cur_n = gaussian_kernel(opts.input_smooth_sigma, gauss_cur);
  else { gauss_cur_n = 1; gauss_cur[0] = 1.f; }
  have_reference = true;
  capacity = 0;   // pointer tables depend on pyr_floats
  if (int e = build_config()) return e;
  hp_main_type = -1;
  if (ic) {
    if (int e = launch_ecc_precompute(cfg, d_hp_trans.as<EccHpCache>(), nullptr, stream)) return e;
  }
  return SSK_OK;
}

int Ecch::build_config() {
  memset(&cfg, 0, sizeof(cfg));
  cfg.nlevels = nlevels;
  for (int l = 0; l < nlevels; ++l) {
    EccLevel &L = cfg.lv[l];
    L.cols = lw[l]; L.rows = lh[l];
    L.ref = ref_pyr.as<float>() + loff[l];
    L.refmask = have_ref_mask ? ref_mask.as<uint8_t>() + loff[l] : nullptr;
    L.gx = ref_gx.p ? ref_gx.as<float>() + loff[l] : nullptr;
    L.gy = ref_gy.p ? ref_gy.as<float>() + loff[l] : nullptr;
    L.cur_off = loff[l];
    L.RMA = rma[l];
  }
  // TMA-staged passes of the inverse-compositional solvers on the large levels (ssk_ecc_impl.cuh::pass_ic_tma): rows must
  // be 16-byte multiples for the copy engine; small levels stay on the gather pass (latency-bound either way)
  cfg.tma_levels = 0;
  cfg.pyr_base = cur_pyr.as<float>();
  cfg.pyr_floats = pyr_floats;
  const bool ic_method = opts.method == SSK_ECC_INVERSE_COMPOSITIONAL || opts.method == SSK_ECC_INVERSE_COMPOSITIONAL_LM;
  static const bool no_tma = getenv("SSK_ECC_NO_TMA") != nullptr;
  if (ic_method && !no_tma && cur_pyr.p && capacity > 0 && ref_gx.p && ref_gy.p) {
    for (int l = 0; l < std::min(nlevels, kTmaLevels); ++l) {
      if ((lw[l] & 3) || (int64_t)lw[l] * lh[l] < 32768) break;
      const int64_t step = (int64_t)lw[l] * 4;
      if (!encode_tmap_3d_f32(cfg.tm[l][0], cur_pyr.as<float>() + loff[l], lw[l], lh[l], capacity, step, pyr_floats * 4, ecc_tma_win_w(), ecc_tma_win_h()) ||
          !encode_tmap_2d_f32(cfg.tm[l][1], ref_pyr.as<float>() + loff[l], lw[l], lh[l], step, ecc_tma_tile_w(), ecc_tma_tile_h()) ||
          !encode_tmap_2d_f32(cfg.tm[l][2], ref_gx.as<float>() + loff[l], lw[l], lh[l], step, ecc_tma_tile_w(), ecc_tma_tile_h()) ||
          !encode_tmap_2d_f32(cfg.tm[l][3], ref_gy.as<float>() + loff[l], lw[l], lh[l], step, ecc_tma_tile_w(), ecc_tma_tile_h()))
        break;
      cfg.tma_levels = l + 1;
    }
  }
  cfg.method = opts.method;
  cfg.interp = remap_interp(opts.interpolation);
  cfg.max_iterations = opts.max_iterations;
  cfg.update_step_scale = opts.update_step_scale;
  cfg.epsx = opts.epsx;
  cfg.max_epse = 1e-4;
  cfg.motion_type = motion_type;
  cfg.translation_first = translation_first;
  cfg.check_rho = check_rho;
  cfg.min_rho = min_rho;
  cfg.final_scale = final_scale;
  cfg.hp_trans = d_hp_trans.as<EccHpCache>();
  cfg.hp_main = d_hp_main.as<EccHpCache>();
  cfg.hp_main_mode = 0;
  cfg.trace = trace_capacity ? d_trace.as<float>() + 4 : nullptr;
  cfg.trace_capacity = trace_capacity;
  cfg.trace_count = trace_capacity ? reinterpret_cast<int *>(d_trace.p) : nullptr;
  return SSK_OK;
}

int Ecch::reserve(int batch) {
  SSK_REQUIRE(have_reference, "c_ecch: reference image must be set first");
  if (batch <= capacity) return SSK_OK;
  const int64_t n0 = (int64_t)lw[0] * lh[0];
  if (int e = cur_pyr.ensure((size_t)batch * pyr_floats * 4)) return e;
  if (int e = src0.ensure((size_t)batch * n0 * 4)) return e;
  if (int e = d_frames.ensure((size_t)batch * sizeof(EccFrame))) return e;
  if (int e = h_frames.ensure((size_t)batch * sizeof(EccFrame))) return e;
  // pointer tables: [level][slot] pyramid pointers, [slot] level-0 source scratch
  std::vector<float *> tab((size_t)nlevels * batch), s0(batch);
  for (int l = 0; l < nlevels; ++l)
    for (int b = 0; b < batch; ++b) tab[(size_t)l * batch + b] = cur_pyr.as<float>() + (int64_t)b * pyr_floats + loff[l];
  for (int b = 0; b < batch; ++b) s0[b] = src0.as<float>() + (int64_t)b * n0;
  if (int e = d_lvl_ptrs.ensure(tab.size() * sizeof(float *))) return e;
  if (int e = d_src0_ptrs.ensure(s0.size() * sizeof(float *))) return e;
  SSK_CUDA(cudaMemcpyAsync(d_lvl_ptrs.p, tab.data(), tab.size() * sizeof(float *), cudaMemcpyHostToDevice, stream));
  SSK_CUDA(cudaMemcpyAsync(d_src0_ptrs.p, s0.data(), s0.size() * sizeof(float *), cudaMemcpyHostToDevice, stream));
  SSK_CUDA(cudaStreamSynchronize(stream));   // tab / s0 are stack-lifetime
  capacity = batch;
  memset(h_frames.p, 0, (size_t)batch * sizeof(EccFrame));
  return SSK_OK;
}

int Ecch::prepare_current(const float *const *d_src_ptrs, int batch) {
  SSK_REQUIRE(batch <= capacity, "c_ecch: batch exceeds reserved capacity");
  float *const *lvl = d_lvl_ptrs.as<float *>();
  SepFilterArgs sf = {};
  sf.rows = lh[0]; sf.cols = lw[0]; sf.batch = batch;
  sf.src_ptrs = d_src_ptrs; sf.dst_ptrs = lvl;
  sf.kxn = sf.kyn = gauss_cur_n;
  memcpy(sf.kx, gauss_cur, sizeof(float) * gauss_cur_n);
  memcpy(sf.ky, gauss_cur, sizeof(float) * gauss_cur_n);
  if (int e = launch_sepfilter(sf, stream)) return e;
  for (int l = 1; l < nlevels; ++l) {
    PyrDownArgs pd = {};
    pd.src.step = (int64_t)lw[l - 1] * 4; pd.src.rows = lh[l - 1]; pd.src.cols = lw[l - 1];
    pd.src.depth = SSK_32F; pd.src.cn = 1; pd.src.scale = 1.f;
    pd.src_ptrs = reinterpret_cast<const void *const *>(lvl + (size_t)(l - 1) * capacity);
    pd.dst_ptrs = lvl + (size_t)l * capacity;
    pd.dst_rows = lh[l]; pd.dst_cols = lw[l]; pd.batch = batch; pd.post_scale = 1.f;
    if (int e = launch_pyrdown(pd, stream)) return e;
  }
  return SSK_OK;
}

int Ecch::prepare_current_mask(const uint8_t *d_mask) {
  SSK_REQUIRE(have_reference && capacity >= 1, "c_ecch: current mask before the current image");
  if (int e = cur_mask.ensure((size_t)pyr_floats)) return e;
  if (int e = cur_mask_tmp.ensure((size_t)lw[0] * lh[0])) return e;
  uint8_t *mp = cur_mask.as<uint8_t>();
  SSK_CUDA(cudaMemcpyAsync(mp, d_mask, (size_t)lw[0] * lh[0], cudaMemcpyDeviceToDevice, stream));
  for (int l = 1; l < nlevels; ++l)
    if (int e = launch_resize_nearest_u8(mp + loff[l - 1], lh[l - 1], lw[l - 1], mp + loff[l], lh[l], lw[l], stream)) return e;
  if (opts.method == SSK_ECC_FORWARD_ADDITIVE) {
    // c_ecc_forward_additive::set_current_image: cv::erode(mask, 5x5, BORDER_REPLICATE) unless the mask is all set
    // (eroding an all-set mask leaves it all set, so the test is not needed)
    for (int l = 0; l < nlevels; ++l) {
      const int n = lw[l] * lh[l];
      SSK_CUDA(cudaMemcpyAsync(cur_mask_tmp.p, mp + loff[l], n, cudaMemcpyDeviceToDevice, stream));
      if (int e = launch_erode5_u8(cur_mask_tmp.as<uint8_t>(), lw[l], mp + loff[l], lw[l], lh[l], lw[l], 1, stream)) return e;
    }
  }
  cur_mask_pending = true;
  return SSK_OK;
}

int Ecch::hp_mode_for_next_align() const {
  const bool ic = opts.method == SSK_ECC_INVERSE_COMPOSITIONAL || opts.method == SSK_ECC_INVERSE_COMPOSITIONAL_LM;
  if (!ic) return 0;
  if (motion_type == SSK_MOTION_TRANSLATION || motion_type == SSK_MOTION_AFFINE) return 0;
  // parameter-dependent steepest-descent images (euclidean, homography):
  //  - with the translation-first pass the solvers see M=2 / M=main alternately and rebuild jac on every
  //    align (jac.size() != M test, ecc2.cc:1733, 1985)  -> per frame
  //  - otherwise jac is built once, by the first align after the reference changed -> capture, then reuse
  if (translation_first) return 1;
  return first_align_pending ? 2 : 0;
}

int Ecch::align(int batch, const ssk_transform &t0) {
  SSK_REQUIRE(have_reference, "c_ecch: reference image must be set first");
  SSK_REQUIRE(batch <= capacity, "c_ecch: batch exceeds reserved capacity");
  motion_type = t0.motion_type;
  if (int e = build_config()) return e;
  const bool ic = opts.method == SSK_ECC_INVERSE_COMPOSITIONAL || opts.method == SSK_ECC_INVERSE_COMPOSITIONAL_LM;
  if (ic && hp_main_type != motion_type) {
    // reference-side jac / Hp of the main transform are (re)built when the parameter count changes (ecc2.cc:1733, 1985)
    const bool param_free = motion_type == SSK_MOTION_TRANSLATION || motion_type == SSK_MOTION_AFFINE;
    if (param_free) {
      if (int e = launch_ecc_precompute(cfg, nullptr, d_hp_main.as<EccHpCache>(), stream)) return e;
    }
    first_align_pending = !param_free;
    hp_main_type = motion_type;
  }
  // frame records are initialised on the device: no host staging, the call stays fully asynchronous
  SSK_REQUIRE(!cur_mask_pending || batch == 1, "c_ecch: a current mask applies to single-frame alignment only");
  const uint8_t *mask_base = cur_mask_pending ? cur_mask.as<uint8_t>() : nullptr;
  cur_mask_pending = false;
  if (int e = launch_ecc_init_frames(device_frames(), batch, t0, cur_pyr.as<float>(), pyr_floats, mask_base, stream)) return e;
  int done = 0;
  const int mode = hp_mode_for_next_align();
  if (mode == 2) {
    cfg.hp_main_mode = 2;
    if (int e = launch_ecc(cfg, device_frames(), 1, cluster_size, stream)) return e;
    first_align_pending = false;
    done = 1;
    cfg.hp_main_mode = 0;
  } else {
    cfg.hp_main_mode = mode;
  }
  // CTAs per frame: 8 for short batches (latency of one frame), 4 once the batch outnumbers the 33 clusters of 8 a B200
  // holds, 2 for batches of several hundred frames (fewer cluster barriers per frame; the tail of the last wave no longer
  // matters).  Measured on config #2, ECC stage per 128 frames (gpurun_out/s5_sweep.log): batch 128: 2.47 / 2.28 / 2.79 ms
  // with 8 / 4 / 2; batch 296: 2.12 (4) / 2.16 (2); batch 512: 2.04 (4) / 1.97 (2); batch 1024: 1.98 (4) / 1.84 (2) / 2.08 (1)
  const int left = batch - done;
  const int cs = cluster_fixed ? cluster_size : (left >= 400 ? 2 : left >= 48 ? 4 : 8);
  if (batch > done)
    if (int e = launch_ecc(cfg, device_frames() + done, batch - done, cs, stream)) return e;
  return SSK_OK;
}

int Ecch::enable_trace(int max_records) {
  trace_capacity = max_records > 0 ? max_records : 0;
  if (!trace_capacity) return SSK_OK;
  if (int e = d_trace.ensure(16 + (size_t)trace_capacity * kTraceRec * 4)) return e;
  SSK_CUDA(cudaMemsetAsync(d_trace.p, 0, 16, stream));
  return SSK_OK;
}

int Ecch::fetch_trace(float *out, int max_records, int *n) {
  *n = 0;
  if (!trace_capacity) return SSK_OK;
  int cnt = 0;
  SSK_CUDA(cudaMemcpyAsync(&cnt, d_trace.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  SSK_CUDA(cudaStreamSynchronize(stream));
  cnt = std::min(cnt, std::min(max_records, trace_capacity));
  if (cnt > 0) SSK_CUDA(cudaMemcpy(out, d_trace.as<float>() + 4, (size_t)cnt * kTraceRec * 4, cudaMemcpyDeviceToHost));
  SSK_CUDA(cudaMemsetAsync(d_trace.p, 0, 16, stream));
  *n = cnt;
  return SSK_OK;
}

int Ecch::download_frames(int batch) {
  SSK_CUDA(cudaMemcpyAsync(h_frames.p, d_frames.p, (size_t)batch * sizeof(EccFrame), cudaMemcpyDeviceToHost, stream));
  return SSK_OK;
}

// ------------------------------------------------------------------------------------------------
Acc::~Acc() {
  if (own_stream && stream) cudaStreamDestroy(stream);
}

int Acc::ensure(int r, int c, int n) {
  if (frames >= 1 || acc.p) {
    SSK_REQUIRE(r == rows && c == cols, "frame accumulation: current frame and accumulator sizes not match");
    SSK_REQUIRE(n == cn, "frame accumulation: current frame and accumulator channel count not match");
    return SSK_OK;
  }
  rows = r; cols = c; cn = n;
  const size_t npix = (size_t)r * c;
  if (kind == SSK_ACC_BAYER_AVERAGE) {
    if (int e = acc.ensure(npix * 3 * 4)) return e;
    if (int e = wacc.ensure(npix * 3 * 4)) return e;
    SSK_CUDA(cudaMemsetAsync(acc.p, 0, npix * 3 * 4, stream));
    SSK_CUDA(cudaMemsetAsync(wacc.p, 0, npix * 3 * 4, stream));
  } else {
    if (int e = acc.ensure(npix * n * 4)) return e;
    if (int e = wacc.ensure(npix * 4)) return e;
    SSK_CUDA(cudaMemsetAsync(acc.p, 0, npix * n * 4, stream));
    SSK_CUDA(cudaMemsetAsync(wacc.p, 0, npix * 4, stream));
  }
  frames = 0;
  return SSK_OK;
}

int Acc::clear() {
  acc.release();
  wacc.release();
  rmap.release();
  rows = cols = cn = 0;
  frames = 0;
  have_map = false;
  rmap_explicit = false;
  return SSK_OK;
}

// ------------------------------------------------------------------------------------------------
Reg::~Reg() {
  if (own_stream && stream) cudaStreamDestroy(stream);
}

int Reg::init(const ssk_registration_options &o, cudaStream_t s, bool own) {
  opts = o;
  stream = s;
  own_stream = own;
  SSK_REQUIRE(o.enable_ecc_registration, "only the ECC registration branch is implemented (enable_ecc_registration)");
  if (int e = make_transform(&default_transform, o.motion_type)) return e;
  current = default_transform;
  ssk_ecch_options eo;
  ssk_ecch_options_default(&eo);
  // c_frame_registration.cc:602-612
  eo.method = o.ecc.ecc_method;
  eo.epsx = o.ecc.eps;
  eo.input_smooth_sigma = o.ecc.input_smooth_sigma;
  eo.reference_smooth_sigma = o.ecc.reference_smooth_sigma;
  eo.update_step_scale = o.ecc.update_step_scale;
  eo.max_iterations = o.ecc.max_iterations;
  eo.maxlevel = o.ecc.ecch_max_level;
  eo.minimum_image_size = o.ecc.ecch_minimum_image_size;
  if (int e = ecch.init(eo, s)) return e;
  ecch.motion_type = o.motion_type;
  // c_frame_registration.cc:815-818
  ecch.translation_first = o.motion_type != SSK_MOTION_TRANSLATION && o.ecc.ecch_estimate_translation_first &&
                           o.ecc.ecch_max_level != 0;
  ecch.check_rho = 1;
  ecch.min_rho = o.ecc.min_rho;
  ecch.final_scale = (o.ecc.scale > 0 && o.ecc.scale != 1) ? 1.0 / o.ecc.scale : 1.0;
  if (o.enable_eccflow_registration) {   // c_frame_registration.cc:637-660
    if (!flowh.p) flowh.p = new (std::nothrow) EccFlow();
    SSK_REQUIRE(flowh.p, "out of memory");
    if (int e = flowh->init(o.eccflow, s)) return e;
  }
  return SSK_OK;
}

// scaleImage (c_frame_registration.cc:230-250): 0.5 -> cv::pyrDown, 1 (or <= 0) -> as is, anything else -> cv::resize(fx = fy =
// scale, INTER_AREA) whose dsize is (cvRound(cols * scale), cvRound(rows * scale))
static bool ecc_scale_is_pyrdown(const ssk_registration_options &o) { return o.ecc.scale > 0 && std::fabs(o.ecc.scale - 0.5) < 1e-2; }
static bool ecc_scale_is_area(const ssk_registration_options &o) { return o.ecc.scale > 0 && o.ecc.scale != 1 && !ecc_scale_is_pyrdown(o); }

static int ecc_image_size(const ssk_registration_options &o, int rows, int cols, int *er, int *ec) {
  if (ecc_scale_is_pyrdown(o)) {
    *er = (rows + 1) / 2; *ec = (cols + 1) / 2;   // cv::pyrDown default dstsize
  } else if (ecc_scale_is_area(o)) {
    SSK_REQUIRE(o.ecc.scale < 1, "ecc.scale > 1 (INTER_AREA up-scaling) is not implemented");
    *er = (int)lrint(rows * o.ecc.scale); *ec = (int)lrint(cols * o.ecc.scale);
    SSK_REQUIRE(*er >= 1 && *ec >= 1, "ecc.scale leaves an empty ECC image");
  } else {
    *er = rows; *ec = cols;
  }
  return SSK_OK;
}

static int scale_to_ecc_image(const ssk_registration_options &o, const Img &geom, const void *const *src_ptrs, float *dst,
                              float *const *dst_ptrs, int er, int ec, int batch, cudaStream_t stream) {
  if (ecc_scale_is_pyrdown(o)) {
    PyrDownArgs pd = {};
    pd.src = geom; pd.src_ptrs = src_ptrs; pd.dst = dst; pd.dst_ptrs = dst_ptrs; pd.dst_rows = er; pd.dst_cols = ec; pd.batch = batch;
    pd.post_scale = 1.f;
    return launch_pyrdown(pd, stream);
  }
  if (ecc_scale_is_area(o)) {
    ResizeAreaArgs ra = {};
    ra.src = geom; ra.src_ptrs = src_ptrs; ra.dst = dst; ra.dst_ptrs = dst_ptrs; ra.dst_rows = er; ra.dst_cols = ec; ra.batch = batch;
    ra.inv_scale_x = ra.inv_scale_y = o.ecc.scale;
    return launch_resize_area(ra, stream);
  }
  return launch_to_gray(geom, src_ptrs, dst, dst_ptrs, batch, stream);
}

int Reg::setup_reference(const Img &frame, const uint8_t *d_mask, int64_t mask_step) {
  SSK_REQUIRE(!opts.ecc.replace_planetary_disk_with_mask, "replace_planetary_disk_with_mask is not implemented");
  ref_rows = frame.rows; ref_cols = frame.cols;
  if (int e = ecc_image_size(opts, frame.rows, frame.cols, &ecc_rows, &ecc_cols)) return e;
  if (int e = staging.ensure((size_t)ecc_rows * ecc_cols * 4)) return e;
  float *d_ecc = staging.as<float>();
  if (int e = scale_to_ecc_image(opts, frame, nullptr, d_ecc, nullptr, ecc_rows, ecc_cols, 1, stream)) return e;
  const uint8_t *d_ecc_mask = nullptr;
  if (d_mask) {
    if (int e = mask_tmp.ensure((size_t)ecc_rows * ecc_cols)) return e;
    SSK_REQUIRE(!ecc_scale_is_area(opts), "reference masks with ecc.scale other than 0.5 / 1 (8-bit INTER_AREA) are not implemented");
    if (ecc_rows != frame.rows) {   // scaleImage: pyrDown(mask) >= 250 (c_frame_registration.cc:237-241)
      if (int e = launch_pyrdown_mask_u8(d_mask, mask_step, frame.rows, frame.cols, mask_tmp.as<uint8_t>(), ecc_rows, ecc_cols, 250, stream)) return e;
    } else {
      SSK_CUDA(cudaMemcpy2DAsync(mask_tmp.p, ecc_cols, d_mask, mask_step, ecc_cols, ecc_rows, cudaMemcpyDeviceToDevice, stream));
    }
    d_ecc_mask = mask_tmp.as<uint8_t>();
  }
  if (normalize_enabled()) {
    if (int e = d_one_ptr.ensure(sizeof(float *))) return e;
    SSK_CUDA(cudaMemcpyAsync(d_one_ptr.p, &d_ecc, sizeof(float *), cudaMemcpyHostToDevice, stream));
    SSK_CUDA(cudaStreamSynchronize(stream));   // &d_ecc is a stack address
    if (int e = normalize(d_one_ptr.as<float *>(), 1, d_ecc_mask)) return e;
  }
  if (int e = ecch.set_reference(d_ecc, ecc_rows, ecc_cols, d_ecc_mask)) return e;
  if (flow_enabled()) {
    // c_frame_registration.cc:637-660: c_eccflow sees the full-size ECC image (create_ecc_image) and the full-size mask
    if (int e = flow_img.ensure((size_t)frame.rows * frame.cols * 4)) return e;
    if (int e = launch_to_gray(frame, nullptr, flow_img.as<float>(), nullptr, 1, stream)) return e;
    const uint8_t *dm = nullptr;
    if (d_mask) {
      if (int e = flow_mask.ensure((size_t)frame.rows * frame.cols)) return e;
      SSK_CUDA(cudaMemcpy2DAsync(flow_mask.p, frame.cols, d_mask, mask_step, frame.cols, frame.rows, cudaMemcpyDeviceToDevice, stream));
      dm = flow_mask.as<uint8_t>();
    }
    if (int e = flowh->set_reference(flow_img.as<float>(), frame.rows, frame.cols, dm)) return e;
  }
  have_current = false;
  return SSK_OK;
}

int Reg::reserve_batch(int batch) {
  SSK_REQUIRE(ecch.have_reference, "c_frame_registration: setup_reference_frame() must be called first");
  return ecch.reserve(batch);
}

int Reg::prepare(const Img &geom, const void *const *d_frame_ptrs, int batch, const uint8_t *d_mask, int64_t mask_step) {
  SSK_REQUIRE(ecch.have_reference, "c_frame_registration: setup_reference_frame() must be called first");
  SSK_REQUIRE(geom.rows == ref_rows && geom.cols == ref_cols, "current frame size differs from the reference frame size");
  if (int e = ecch.reserve(batch)) return e;
  const bool scaled = ecc_images_ready;
  ecc_images_ready = false;
  if (scaled) SSK_REQUIRE(!flow_enabled(), "internal: pre-scaled ECC images with eccflow (needs the full-size frame)");
  else if (int e = scale_to_ecc_image(opts, geom, d_frame_ptrs, nullptr, ecch.level0_scratch_ptrs(), ecc_rows, ecc_cols, batch, stream)) return e;
  const uint8_t *d_ecc_mask = nullptr;
  if (d_mask) {
    // scaleImage of the current mask (c_frame_registration.cc:230-250): pyrDown(mask) >= 250 for ecc.scale 0.5, as is for 1
    SSK_REQUIRE(batch == 1, "c_frame_registration: a current mask applies to single-frame registration");
    SSK_REQUIRE(!ecc_scale_is_area(opts), "current masks with ecc.scale other than 0.5 / 1 (8-bit INTER_AREA) are not implemented");
    if (int e = mask_tmp.ensure((size_t)ecc_rows * ecc_cols)) return e;
    if (ecc_rows != geom.rows) {
      if (int e = launch_pyrdown_mask_u8(d_mask, mask_step, geom.rows, geom.cols, mask_tmp.as<uint8_t>(), ecc_rows, ecc_cols, 250, stream)) return e;
    } else {
      SSK_CUDA(cudaMemcpy2DAsync(mask_tmp.p, ecc_cols, d_mask, mask_step, ecc_cols, ecc_rows, cudaMemcpyDeviceToDevice, stream));
    }
    d_ecc_mask = mask_tmp.as<uint8_t>();
  }
  if (normalize_enabled()) {
    if (int e = normalize(ecch.level0_scratch_ptrs(), batch, d_ecc_mask)) return e;
  }
  if (int e = ecch.prepare_current(ecch.level0_scratch_ptrs(), batch)) return e;
  if (d_ecc_mask)
    if (int e = ecch.prepare_current_mask(d_ecc_mask)) return e;
  if (flow_enabled()) {
    // c_frame_registration.cc:773-785, 900-917: the full-size ECC image and mask of the current frame
    if (int e = flowh->reserve(batch)) return e;
    if (int e = launch_to_gray(geom, d_frame_ptrs, nullptr, flowh->level0_ptrs(), batch, stream)) return e;
    const uint8_t *dm = nullptr;
    if (d_mask) {
      if (int e = flow_mask.ensure((size_t)geom.rows * geom.cols)) return e;
      SSK_CUDA(cudaMemcpy2DAsync(flow_mask.p, geom.cols, d_mask, mask_step, geom.cols, geom.rows, cudaMemcpyDeviceToDevice, stream));
      dm = flow_mask.as<uint8_t>();
    }
    if (int e = flowh->build_current(batch, dm)) return e;
  }
  return SSK_OK;
}

// ecc_normalize (ecc2.cc:385-397): dst = src - ecc_upscale(ecc_downscale(src, level, BORDER_REPLICATE), src.size()),
// zero under the mask.  The pyrDown chain of every frame lives in norm_buf; pyrUp results overwrite the chain on the
// way back up and the last pyrUp is fused with the subtraction.
int Reg::normalize(float *const *d_img_ptrs, int batch, const uint8_t *d_mask) {
  const int L = opts.ecc.normalization_scale;
  SSK_REQUIRE(L >= 1 && L <= 12, "ecc_normalize: normalization_scale 1..12");
  int cw[16], ch[16];
  int64_t off[16], total = 0;
  cw[0] = ecc_cols; ch[0] = ecc_rows;
  for (int k = 1; k <= L; ++k) {
    cw[k] = (cw[k - 1] + 1) / 2; ch[k] = (ch[k - 1] + 1) / 2;
    SSK_REQUIRE(std::min(cw[k], ch[k]) >= 2, "ecc_normalize: normalization_scale too deep for this image size");
    off[k] = total;
    total += ((int64_t)cw[k] * ch[k] + 63) & ~(int64_t)63;
  }
  if (batch > norm_capacity) {
    if (int e = norm_buf.ensure((size_t)batch * total * 4)) return e;
    std::vector<float *> tab((size_t)(L + 1) * batch);
    for (int k = 1; k <= L; ++k)
      for (int b = 0; b < batch; ++b) tab[(size_t)k * batch + b] = norm_buf.as<float>() + (int64_t)b * total + off[k];
    if (int e = norm_ptrs.ensure(tab.size() * sizeof(float *))) return e;
    SSK_CUDA(cudaStreamSynchronize(stream));
    SSK_CUDA(cudaMemcpy(norm_ptrs.p, tab.data(), tab.size() * sizeof(float *), cudaMemcpyHostToDevice));
    norm_capacity = batch;
  }
  float *const *lv = norm_ptrs.as<float *>();
  const int cap = norm_capacity;
  for (int k = 1; k <= L; ++k) {
    PyrDownArgs pd = {};
    pd.src.step = (int64_t)cw[k - 1] * 4; pd.src.rows = ch[k - 1]; pd.src.cols = cw[k - 1];
    pd.src.depth = SSK_32F; pd.src.cn = 1; pd.src.scale = 1.f;
    pd.src_ptrs = reinterpret_cast<const void *const *>(k == 1 ? d_img_ptrs : lv + (size_t)(k - 1) * cap);
    pd.dst_ptrs = lv + (size_t)k * cap; pd.dst_rows = ch[k]; pd.dst_cols = cw[k]; pd.batch = batch; pd.post_scale = 1.f;
    pd.border = SSK_BORDER_REPLICATE;
    if (int e = launch_pyrdown(pd, stream)) return e;
  }
  for (int k = L; k >= 1; --k) {
    PyrUpArgs pu = {};
    pu.src_ptrs = lv + (size_t)k * cap; pu.rows = ch[k]; pu.cols = cw[k];
    pu.dst_rows = ch[k - 1]; pu.dst_cols = cw[k - 1]; pu.batch = batch;
    if (k == 1) { pu.dst_ptrs = d_img_ptrs; pu.minuend_ptrs = d_img_ptrs; pu.mask = d_mask; }
    else pu.dst_ptrs = lv + (size_t)(k - 1) * cap;
    if (int e = launch_pyrup(pu, stream)) return e;
  }
  return SSK_OK;
}

int Reg::register_batch(int batch) {
  // c_frame_registration.cc:744: every frame restarts from the default parameters
  if (int e = ecch.align(batch, default_transform)) return e;
  // c_frame_registration.cc:900-917: _eccflow.compute(ecc_image, _current_remap, eccflow_mask) for the registered frames
  if (flow_enabled()) return flowh->compute(batch, ecch.device_frames(), nullptr);
  return SSK_OK;
}

}  // namespace ssk
