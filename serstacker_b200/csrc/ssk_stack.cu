// ssk_stack_*: the per-frame loop of c_image_stacking_pipeline::process_input_sequence
// (core/pipeline/c_image_stacking_pipeline/c_image_stacking_pipeline.cc:1358-1862) for batches of frames:
//   compute_weights (:1466, 2013-2020) -> register_frame (:1539) -> custom_remap(frame, mask) (:1644-1651)
//   -> custom_remap(weights, BORDER_CONSTANT) (:1653-1660) -> weights *= mask/255 (:1704-1714) -> add (:1753-1779)
// Every stage is one batched kernel launch; frames never leave the device and the call is asynchronous.
#include <cmath>
#include <cstring>
#include <new>
#include <cstdlib>
#include "ssk_engine.cuh"
#include "ssk_eccflow.cuh"
#include "ssk_nvtx.h"
#include "ssk_upscale.cuh"

using namespace ssk;

namespace {

__global__ void k_fill_jobs(const EccFrame *reg, const void *const *frame_ptrs, float *const *weight_ptrs, const double *w1_stats,
                            FrameJob *jobs, int n, int *accumulated, int have_reg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FrameJob j;
  j.frame = frame_ptrs[i];
  j.weights = weight_ptrs ? weight_ptrs[i] : nullptr;
  // compute_local_variance_map releases the map for a flat frame (sum of gradients == 0): the pipeline then
  // accumulates that frame with the plain 8U mask (current_weights.empty(), c_image_stacking_pipeline.cc:1704)
  if (w1_stats && !(w1_stats[i * 4] > 0)) j.weights = nullptr;
  if (have_reg) {
    j.map = reg[i].map;
    j.ok = reg[i].ok;
  } else {
    for (int k = 0; k < 9; ++k) j.map.c[k] = 0.f;
    j.map.type = MAP_TRANSLATION;
    j.ok = 1;
  }
  j.pad = 0;
  jobs[i] = j;
  if (j.ok) atomicAdd(accumulated, 1);
}

constexpr int kRing = 4;

}  // namespace

struct ssk_stack {
  ssk_stack_options o;
  cudaStream_t stream = nullptr;
  ssk_reg reg_h;
  ssk_acc acc_h;
  Tables tab;
  int max_batch = 0;
  // geometry of the sequence (fixed by set_reference)
  int rows = 0, cols = 0, type = -1, bpp = 0;
  int upscale = 0, out_rows = 0, out_cols = 0;   // frame_upscale_after_align: option in force and the accumulator size
  bool have_reference = false;
  // per-slot device buffers
  DevBuf frame_slots, weight_slots, half_slots, half2_slots, gmap_slots, partials, stats, jobs, counter;
  DevBuf d_slot_ptrs, d_weight_ptrs, d_half_ptrs, d_half2_ptrs, d_gmap_ptrs;
  DevBuf gmap2_slots, d_gmap2_ptrs;      // sharpness_measure.uscale > 0: INTER_AREA-reduced maps
  DevBuf d_user_ptrs[kRing];
  PinnedBuf h_user_ptrs[kRing];
  // TMA staging of the fused kernel (32F single-channel frames): 128-byte tensor maps per frame slot / weight slot,
  // and per user frame of a device-frame chunk (same ring as the pointer tables)
  DevBuf d_tmaps_slots, d_tmaps_weights, d_tmaps_user[kRing];
  PinnedBuf h_tmaps_user[kRing];
  bool tma_slots = false, tma_weights = false;
  cudaEvent_t ring_ev[kRing] = {};
  int ring_pos = 0;
  PinnedBuf h_counter;
  int w1_rows = 0, w1_cols = 0, w1_nb = 0;
  bool frames_aligned = true;            // every frame pointer of the current chunk is 16-byte aligned
  cudaEvent_t ev[5] = {};
  cudaStream_t side = nullptr;           // border-ring kernel of the fused warp+accumulate stage
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // Device-frame calls that span several chunks let the ring kernel of chunk k run behind the registration of chunk
  // k + 1 (joined at the end of the call): what the ring kernel reads exists twice, chunks alternate between the copies.
  DevBuf weight_slots_b, d_weight_ptrs_b, d_tmaps_weights_b, jobs_b;
  cudaEvent_t ev_join_b = nullptr;
  int parity = 0;
  bool join_recorded[2] = {false, false};
  int ring_pending = -1;                 // copy whose ring kernel the handle's stream has not been made to wait for yet
  cudaEvent_t ev_ring_end = nullptr;     // (timing) end of the last chunk's ring kernel, on the side stream
  DevBuf ref_staging, axis_tab;
  // bayer_average: demosaiced (BGR, frame depth) copies of the chunk's raw frames feed the registration; the raw frames
  // feed the accumulator (c_image_stacking_pipeline_base.cc:221-236, c_image_stacking_pipeline.cc:1730-1752)
  DevBuf bgr_slots, d_bgr_ptrs, ref_bgr;
  int ref_cn = 0;                        // channels of the reference frame as given
  int axis_tab_built = 0;                // W1 up-sampling tables of this geometry are in axis_tab
  // host frames: sub-chunks of `host_chunk` frames rotate through `nsets` sets of frame slots; a copy stream uploads
  // sub-chunk k+1 while sub-chunk k is processed
  cudaStream_t copy_stream = nullptr;
  static constexpr int kMaxSets = 8;
  cudaEvent_t set_free[kMaxSets] = {}, set_full[kMaxSets] = {};
  int host_chunk = 0, nsets = 1, set_pos = 0;
  // registration records of the last kRecRing chunks (chunk `ticket` lives in slot ticket % kRecRing)
  static constexpr int kRecRing = 4;
  DevBuf rec_all;
  PinnedBuf h_rec_all;
  cudaEvent_t rec_ev[kRecRing] = {};
  // device-frame chunks leave their border-ring kernel running on the side stream: it reads the caller's frames, so
  // ssk_stack_wait(ticket) also waits for the chunk's ring kernel (recorded on the side stream)
  cudaEvent_t rec_ring_ev[kRecRing] = {};
  bool rec_ring_valid[kRecRing] = {};
  int rec_n[kRecRing] = {};
  int64_t rec_ticket[kRecRing] = {-1, -1, -1, -1};
  int64_t next_ticket = 0;
  ~ssk_stack() {
    for (auto &e : rec_ev) if (e) cudaEventDestroy(e);
    for (auto &e : rec_ring_ev) if (e) cudaEventDestroy(e);
    for (auto &e : set_free) if (e) cudaEventDestroy(e);
    for (auto &e : set_full) if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (auto &e : ring_ev) if (e) cudaEventDestroy(e);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_join_b) cudaEventDestroy(ev_join_b);
    if (ev_ring_end) cudaEventDestroy(ev_ring_end);
    if (side) cudaStreamDestroy(side);
    if (stream) cudaStreamDestroy(stream);
  }
};

static int stack_alloc_slots(ssk_stack *h) {
  const int B = h->max_batch;
  const int d = type_depth(h->type), cn = type_cn(h->type);
  const size_t frame_bytes = (size_t)h->rows * h->cols * cn * depth_bytes(d);
  const size_t npix = (size_t)h->rows * h->cols;
  std::vector<void *> p(B);
  if (int e = h->frame_slots.ensure(frame_bytes * B)) return e;
  for (int b = 0; b < B; ++b) p[b] = h->frame_slots.as<char>() + frame_bytes * b;
  if (int e = h->d_slot_ptrs.ensure(sizeof(void *) * B)) return e;
  SSK_CUDA(cudaMemcpy(h->d_slot_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
  h->tma_slots = false;
  if (d == SSK_32F && cn == 1 && !getenv("SSK_NO_TMA")) {
    std::vector<unsigned char> maps((size_t)B * 128);
    bool ok = true;
    for (int b = 0; b < B && ok; ++b)
      ok = encode_tmap_2d_f32(maps.data() + (size_t)b * 128, p[b], h->cols, h->rows, (int64_t)h->cols * 4, staged_box_w(), staged_box_h());
    if (ok) {
      if (int e = h->d_tmaps_slots.ensure(maps.size())) return e;
      SSK_CUDA(cudaMemcpy(h->d_tmaps_slots.p, maps.data(), maps.size(), cudaMemcpyHostToDevice));
      for (int r = 0; r < kRing; ++r) {
        if (int e = h->d_tmaps_user[r].ensure(maps.size())) return e;
        if (int e = h->h_tmaps_user[r].ensure(maps.size())) return e;
      }
      h->tma_slots = true;
    }
  }
  if (h->o.accumulation_method == SSK_STACK_BAYER_AVERAGE && h->o.enable_registration) {
    const size_t bgr_bytes = frame_bytes * 3;
    if (int e = h->bgr_slots.ensure(bgr_bytes * B)) return e;
    for (int b = 0; b < B; ++b) p[b] = h->bgr_slots.as<char>() + bgr_bytes * b;
    if (int e = h->d_bgr_ptrs.ensure(sizeof(void *) * B)) return e;
    SSK_CUDA(cudaMemcpy(h->d_bgr_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
  }
  if (int e = h->jobs.ensure(sizeof(FrameJob) * B)) return e;
  if (int e = h->jobs_b.ensure(sizeof(FrameJob) * B)) return e;
  if (int e = h->counter.ensure(sizeof(int))) return e;
  if (int e = h->h_counter.ensure(sizeof(int))) return e;
  SSK_CUDA(cudaMemset(h->counter.p, 0, sizeof(int)));
  for (int r = 0; r < kRing; ++r) {
    if (int e = h->d_user_ptrs[r].ensure(sizeof(void *) * B)) return e;
    if (int e = h->h_user_ptrs[r].ensure(sizeof(void *) * B)) return e;
    if (!h->ring_ev[r]) SSK_CUDA(cudaEventCreateWithFlags(&h->ring_ev[r], cudaEventDisableTiming));
  }
  for (auto &e : h->ev) if (!e) SSK_CUDA(cudaEventCreate(&e));
  if (int e = h->rec_all.ensure(sizeof(EccFrame) * B * ssk_stack::kRecRing)) return e;
  if (int e = h->h_rec_all.ensure(sizeof(EccFrame) * B * ssk_stack::kRecRing)) return e;
  for (auto &e : h->rec_ev) if (!e) SSK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : h->rec_ring_ev) if (!e) SSK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  // sub-chunk of host frames: large enough to keep the kernels efficient, small enough for >= 2 sets in flight
  h->host_chunk = B >= 32 ? std::max(16, B / 4) : B;
  if (const char *e = getenv("SSK_HOST_CHUNK")) { const int c = atoi(e); if (c >= 1 && c <= B) h->host_chunk = c; }
  h->nsets = std::max(1, std::min<int>(ssk_stack::kMaxSets, B / h->host_chunk));
  if (h->nsets < 2) h->host_chunk = B;   // a single slot set cannot take the next sub-chunk's upload while it is in use
  if (!h->copy_stream) {
    SSK_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto &e : h->set_free) SSK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : h->set_full) SSK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  if (!h->side) {
    SSK_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    SSK_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    SSK_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    SSK_CUDA(cudaEventCreateWithFlags(&h->ev_join_b, cudaEventDisableTiming));
    SSK_CUDA(cudaEventCreate(&h->ev_ring_end));
  }
  if (h->o.accumulation_method == SSK_STACK_WEIGHTED_AVERAGE && h->o.sm_kradius > 0) {
    // pdownscale chain sizes (c_local_variance_sharpness_measure.cc:28-52)
    int r = h->rows, c = h->cols;
    for (int l = 0; l < h->o.sm_dscale && std::min(r, c) >= 4; ++l) { r = (r + 1) / 2; c = (c + 1) / 2; }
    h->w1_rows = r; h->w1_cols = c; h->w1_nb = w1_num_blocks(r, c);
    const size_t half_px = (size_t)((h->rows + 1) / 2) * ((h->cols + 1) / 2);
    if (int e = h->weight_slots.ensure(npix * 4 * B)) return e;
    if (int e = h->half_slots.ensure(half_px * 4 * B)) return e;
    if (int e = h->half2_slots.ensure(half_px * 4 * B)) return e;
    if (int e = h->gmap_slots.ensure((size_t)r * c * 4 * B)) return e;
    if (int e = h->partials.ensure((size_t)B * 2 * h->w1_nb * 8)) return e;
    if (int e = h->stats.ensure((size_t)B * 4 * 8)) return e;
    if (int e = h->axis_tab.ensure((size_t)(h->rows + h->cols + 8) * sizeof(int2))) return e;
    h->axis_tab_built = 0;
    if (int e = h->d_weight_ptrs.ensure(sizeof(void *) * B)) return e;
    if (int e = h->d_half_ptrs.ensure(sizeof(void *) * B)) return e;
    if (int e = h->d_half2_ptrs.ensure(sizeof(void *) * B)) return e;
    if (int e = h->d_gmap_ptrs.ensure(sizeof(void *) * B)) return e;
    for (int b = 0; b < B; ++b) p[b] = h->weight_slots.as<float>() + npix * b;
    SSK_CUDA(cudaMemcpy(h->d_weight_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
    if (int e = h->weight_slots_b.ensure(npix * 4 * B)) return e;
    if (int e = h->d_weight_ptrs_b.ensure(sizeof(void *) * B)) return e;
    std::vector<void *> pb(B);
    for (int b = 0; b < B; ++b) pb[b] = h->weight_slots_b.as<float>() + npix * b;
    SSK_CUDA(cudaMemcpy(h->d_weight_ptrs_b.p, pb.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
    h->tma_weights = false;
    if (!getenv("SSK_NO_TMA")) {
      std::vector<unsigned char> maps((size_t)B * 128), maps_b((size_t)B * 128);
      bool ok = true;
      for (int b = 0; b < B && ok; ++b)
        ok = encode_tmap_2d_f32(maps.data() + (size_t)b * 128, p[b], h->cols, h->rows, (int64_t)h->cols * 4, staged_box_w(), staged_box_h()) &&
             encode_tmap_2d_f32(maps_b.data() + (size_t)b * 128, pb[b], h->cols, h->rows, (int64_t)h->cols * 4, staged_box_w(), staged_box_h());
      if (ok) {
        if (int e = h->d_tmaps_weights.ensure(maps.size())) return e;
        if (int e = h->d_tmaps_weights_b.ensure(maps.size())) return e;
        SSK_CUDA(cudaMemcpy(h->d_tmaps_weights.p, maps.data(), maps.size(), cudaMemcpyHostToDevice));
        SSK_CUDA(cudaMemcpy(h->d_tmaps_weights_b.p, maps_b.data(), maps_b.size(), cudaMemcpyHostToDevice));
        h->tma_weights = true;
      }
    }
    for (int b = 0; b < B; ++b) p[b] = h->half_slots.as<float>() + half_px * b;
    SSK_CUDA(cudaMemcpy(h->d_half_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
    for (int b = 0; b < B; ++b) p[b] = h->half2_slots.as<float>() + half_px * b;
    SSK_CUDA(cudaMemcpy(h->d_half2_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
    for (int b = 0; b < B; ++b) p[b] = h->gmap_slots.as<float>() + (size_t)r * c * b;
    SSK_CUDA(cudaMemcpy(h->d_gmap_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
    if (h->o.sm_uscale > 0) {
      if (int e = h->gmap2_slots.ensure((size_t)r * c * 4 * B)) return e;
      if (int e = h->d_gmap2_ptrs.ensure(sizeof(void *) * B)) return e;
      for (int b = 0; b < B; ++b) p[b] = h->gmap2_slots.as<float>() + (size_t)r * c * b;
      SSK_CUDA(cudaMemcpy(h->d_gmap2_ptrs.p, p.data(), sizeof(void *) * B, cudaMemcpyHostToDevice));
    }
  }
  return SSK_OK;
}

// parts of the handle the multi-GPU combine works on (ssk_multi.cu)
int stack_reduce_parts(ssk_stack *h, cudaStream_t *stream, ssk_acc **acc, int **d_counter) {
  SSK_REQUIRE(h && h->have_reference, "ssk_stack_reduce: set_reference must be called first");
  *stream = h->stream; *acc = &h->acc_h; *d_counter = h->counter.as<int>();
  return SSK_OK;
}

extern "C" {

void ssk_stack_options_default(ssk_stack_options *o) {
  memset(o, 0, sizeof(*o));
  ssk_registration_options_default(&o->registration);
  o->registration.enable_ecc_registration = 1;
  o->accumulation_method = SSK_STACK_AVERAGE;
  o->sm_dscale = 1; o->sm_kradius = 1; o->sm_uscale = 0;   // c_image_stacking_pipeline.h:94-99
  o->enable_registration = 1;
  o->bayer_colorid = SSK_COLORID_BAYER_RGGB;
  o->max_batch = 32;
}

int ssk_stack_create(const ssk_stack_options *opts, ssk_stack **out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return SSK_ERR_CUDA;
  }
  SSK_REQUIRE(opts && out, "ssk_stack_create: null argument");
  SSK_REQUIRE(opts->max_batch >= 1 && opts->max_batch <= 4096, "max_batch 1..4096");
  SSK_REQUIRE(opts->accumulation_method == SSK_STACK_AVERAGE || opts->accumulation_method == SSK_STACK_WEIGHTED_AVERAGE ||
              opts->accumulation_method == SSK_STACK_BAYER_AVERAGE, "ssk_stack: accumulation_method must be average, weighted_average or bayer_average");
  if (opts->accumulation_method == SSK_STACK_BAYER_AVERAGE)
    SSK_REQUIRE(opts->bayer_colorid >= SSK_COLORID_BAYER_RGGB && opts->bayer_colorid <= SSK_COLORID_BAYER_BGGR,
                "ssk_stack: bayer_average needs bayer_colorid RGGB / GRBG / GBRG / BGGR");
  SSK_REQUIRE(opts->sm_uscale >= 0 && opts->sm_uscale <= 12, "sharpness_measure.uscale 0..12");
  SSK_REQUIRE(opts->upscale_option >= SSK_UPSCALE_NONE && opts->upscale_option <= SSK_UPSCALE_X30, "ssk_stack: upscale_option must be none / x2.0 / x1.5 / x3.0");
  if (opts->upscale_option != SSK_UPSCALE_NONE) {
    SSK_REQUIRE(opts->upscale_stage == SSK_UPSCALE_AFTER_ALIGN || opts->upscale_stage == 0,
                "ssk_stack: only frame_upscale_after_align is fused into the loop (up-scale the frames with ssk_upscale_image for before_align)");
    SSK_REQUIRE(opts->accumulation_method != SSK_STACK_BAYER_AVERAGE, "ssk_stack: up-scaling with bayer_average: the reference hands the un-scaled current_remap and the up-scaled mask to c_bayer_average::add, which rejects the size mismatch (c_image_stacking_pipeline.cc:1633-1642, 1750-1751)");
  }
  ssk_stack *h = new (std::nothrow) ssk_stack();
  SSK_REQUIRE(h, "out of memory");
  h->o = *opts;
  h->max_batch = opts->max_batch;
  cudaError_t ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) { delete h; return cuda_fail(ce, "cudaStreamCreate", __FILE__, __LINE__); }
  if (opts->enable_registration) {
    if (int e = h->reg_h.r.init(opts->registration, h->stream, false)) { delete h; return e; }
  }
  h->acc_h.a.kind = opts->accumulation_method == SSK_STACK_BAYER_AVERAGE ? SSK_ACC_BAYER_AVERAGE : SSK_ACC_WEIGHTED_AVERAGE;
  h->acc_h.a.colorid = opts->bayer_colorid;
  h->acc_h.a.stream = h->stream;
  if (int e = get_tables(&h->tab)) { delete h; return e; }
  *out = h;
  return SSK_OK;
}

int ssk_stack_destroy(ssk_stack *h) { delete h; return SSK_OK; }

static int stack_check_mat(const ssk_mat *m, const char *what) {
  if (!m || !m->data || m->rows <= 0 || m->cols <= 0) { set_error(std::string(what) + ": empty image"); return SSK_ERR_INVALID; }
  const int d = type_depth(m->type), cn = type_cn(m->type);
  if (!depth_bytes(d) || cn < 1 || cn > 4) { set_error(std::string(what) + ": unsupported type"); return SSK_ERR_INVALID; }
  if (m->step < (int64_t)m->cols * cn * depth_bytes(d)) { set_error(std::string(what) + ": step smaller than a row"); return SSK_ERR_INVALID; }
  return SSK_OK;
}

int ssk_stack_set_reference(ssk_stack *h, const ssk_mat *image, const ssk_mat *mask, int bpp) {
  SSK_REQUIRE(h && image, "ssk_stack_set_reference: null argument");
  if (int e = stack_check_mat(image, "ssk_stack_set_reference: reference frame")) return e;
  const int d = type_depth(image->type), cn = type_cn(image->type);
  const bool bayer = h->o.accumulation_method == SSK_STACK_BAYER_AVERAGE;
  SSK_REQUIRE(cn == 1 || cn == 3, "frames must be 8U/16U/32F with 1 or 3 channels");
  if (bayer) SSK_REQUIRE(!(image->rows & 1) && !(image->cols & 1), "bayer_average: frame size must be even");
  h->rows = image->rows; h->cols = image->cols; h->bpp = bpp; h->ref_cn = cn;
  // frames are expected in the reference frame's type (a raw Bayer frame per reference pixel for bayer_average); the first
  // frames of another depth re-latch it (a CV_32F master frame over 16-bit input frames), see stack_check_frames
  h->type = bayer ? SSK_MAKETYPE(d, 1) : image->type;
  if (h->o.enable_registration) {
    Img im;
    im.rows = image->rows; im.cols = image->cols; im.depth = d; im.cn = cn; im.scale = bpp_scale(d, bpp);
    const size_t rowb = (size_t)image->cols * cn * depth_bytes(d);
    if (image->mem == SSK_MEM_DEVICE) { im.data = image->data; im.step = image->step; }
    else {
      if (int e = h->ref_staging.ensure(rowb * image->rows)) return e;
      SSK_CUDA(cudaMemcpy2DAsync(h->ref_staging.p, rowb, image->data, image->step, rowb, image->rows, cudaMemcpyHostToDevice, h->stream));
      im.data = h->ref_staging.p; im.step = (int64_t)rowb;
    }
    if (bayer && cn == 1) {
      // a raw Bayer reference is demosaiced as read_input_frame does (c_image_stacking_pipeline_base.cc:221-236)
      const size_t bstep = (size_t)image->cols * 3 * depth_bytes(d);
      if (int e = h->ref_bgr.ensure(bstep * image->rows)) return e;
      if (int e = launch_debayer_nn2(im.data, im.step, d, image->rows, image->cols, h->o.bayer_colorid, h->ref_bgr.p, (int64_t)bstep, h->stream)) return e;
      im.data = h->ref_bgr.p; im.step = (int64_t)bstep; im.cn = 3;
    }
    const uint8_t *d_mask = nullptr;
    int64_t mstep = 0;
    if (mask) { if (int e = mask_to_device(mask, image->rows, image->cols, h->reg_h.st_mask, h->stream, &d_mask, &mstep)) return e; }
    if (int e = h->reg_h.r.setup_reference(im, d_mask, mstep)) return e;
    if (int e = h->reg_h.r.ecch.reserve(h->max_batch)) return e;
    if (h->reg_h.r.flow_enabled())
      if (int e = h->reg_h.r.flowh->reserve(h->max_batch)) return e;
  }
  if (int e = stack_alloc_slots(h)) return e;
  // frame_upscale_after_align: the frames are stacked at the up-scaled size (never during the master-frame pass,
  // c_image_stacking_pipeline.cc:2093-2101; only behind a registration, :1633-1642)
  h->upscale = (h->o.enable_registration && !h->o.generating_master_frame) ? h->o.upscale_option : SSK_UPSCALE_NONE;
  upscale_size(h->upscale, h->cols, h->rows, &h->out_cols, &h->out_rows);
  if (int e = h->acc_h.a.ensure(h->out_rows, h->out_cols, bayer ? 3 : cn)) return e;
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  h->have_reference = true;
  return SSK_OK;
}

// Uploads `n` host frames into slot set `set` on the copy stream (after the set's previous user has finished).
static int stack_upload(ssk_stack *h, const ssk_mat *frames, int n, int set) {
  const int d = type_depth(h->type), cn = type_cn(h->type);
  const size_t rowb = (size_t)h->cols * cn * depth_bytes(d);
  SSK_CUDA(cudaStreamWaitEvent(h->copy_stream, h->set_free[set], 0));
  char *base = h->frame_slots.as<char>() + rowb * h->rows * (size_t)set * h->host_chunk;
  for (int i = 0; i < n; ++i) {
    SSK_REQUIRE(frames[i].mem == SSK_MEM_HOST, "frames of a call must share memory space");
    SSK_CUDA(cudaMemcpy2DAsync(base + rowb * h->rows * i, rowb, frames[i].data, frames[i].step, rowb, h->rows,
                               cudaMemcpyHostToDevice, h->copy_stream));
  }
  SSK_CUDA(cudaEventRecord(h->set_full[set], h->copy_stream));
  return SSK_OK;
}

// One chunk of frames through the per-frame loop.  set >= 0: host frames already being uploaded into slot set `set`;
// set < 0: device frames.  The registration records land in rec_all[rec_off ...].
static int stack_process_chunk(ssk_stack *h, const ssk_mat *frames, int n, int set, int rec_off, bool last_in_call) {
  const int d = type_depth(h->type), cn = type_cn(h->type);
  const size_t rowb = (size_t)h->cols * cn * depth_bytes(d);
  cudaStream_t s = h->stream;
  NvtxStage nvtx;
  nvtx.next("ssk_stack: frame pointers");
  // copy (0 / 1) of the buffers the ring kernel reads; host-frame chunks join their ring kernel at once and stay on copy 0
  (void)last_in_call;
  const bool defer = set < 0 && h->side && !getenv("SSK_NO_SIDE_STREAM") && !getenv("SSK_NO_DEFERRED_RING");
  const int par = set < 0 ? h->parity : 0;
  if (set < 0) h->parity ^= 1;
  float *const *weight_ptrs = (par ? h->d_weight_ptrs_b : h->d_weight_ptrs).as<float *>();
  FrameJob *jobs = (par ? h->jobs_b : h->jobs).as<FrameJob>();
  cudaEvent_t ev_join = par ? h->ev_join_b : h->ev_join;
  if (h->join_recorded[par]) SSK_CUDA(cudaStreamWaitEvent(s, ev_join, 0));   // ring kernel of the chunk that used this copy last
  Img geom;
  geom.rows = h->rows; geom.cols = h->cols; geom.depth = d; geom.cn = cn; geom.scale = bpp_scale(d, h->bpp);
  geom.data = nullptr;
  const void *const *d_frame_ptrs;
  const void *d_tmaps = nullptr;          // tensor maps of this chunk's frames (TMA staging), if available
  h->frames_aligned = true;
  if (frames[0].mem == SSK_MEM_DEVICE) {
    // frames already resident: upload this chunk's pointer table through a small pinned ring
    const int r = h->ring_pos;
    h->ring_pos = (h->ring_pos + 1) % kRing;
    SSK_CUDA(cudaEventSynchronize(h->ring_ev[r]));
    void **hp = h->h_user_ptrs[r].as<void *>();
    for (int i = 0; i < n; ++i) {
      SSK_REQUIRE(frames[i].mem == SSK_MEM_DEVICE && frames[i].step == frames[0].step, "frames of a call must share memory space and step");
      hp[i] = frames[i].data;
      if (reinterpret_cast<uintptr_t>(frames[i].data) & 15) h->frames_aligned = false;
    }
    SSK_CUDA(cudaMemcpyAsync(h->d_user_ptrs[r].p, hp, sizeof(void *) * n, cudaMemcpyHostToDevice, s));
    const bool need_wmaps = h->o.accumulation_method == SSK_STACK_WEIGHTED_AVERAGE && h->o.sm_kradius > 0;
    if (h->tma_slots && (h->tma_weights || !need_wmaps) && h->frames_aligned && !getenv("SSK_NO_TMA")) {
      unsigned char *hm = h->h_tmaps_user[r].as<unsigned char>();
      bool ok = true;
      for (int i = 0; i < n && ok; ++i)
        ok = encode_tmap_2d_f32(hm + (size_t)i * 128, frames[i].data, h->cols, h->rows, (int64_t)frames[i].step, staged_box_w(), staged_box_h());
      if (ok) {
        SSK_CUDA(cudaMemcpyAsync(h->d_tmaps_user[r].p, hm, (size_t)n * 128, cudaMemcpyHostToDevice, s));
        d_tmaps = h->d_tmaps_user[r].p;
      }
    }
    SSK_CUDA(cudaEventRecord(h->ring_ev[r], s));
    d_frame_ptrs = h->d_user_ptrs[r].as<const void *>();
    geom.step = frames[0].step;
  } else {
    SSK_REQUIRE(set >= 0, "internal: host frames without a slot set");
    SSK_CUDA(cudaStreamWaitEvent(s, h->set_full[set], 0));
    d_frame_ptrs = h->d_slot_ptrs.as<const void *>() + (size_t)set * h->host_chunk;
    const bool need_wmaps = h->o.accumulation_method == SSK_STACK_WEIGHTED_AVERAGE && h->o.sm_kradius > 0;
    if (h->tma_slots && (h->tma_weights || !need_wmaps)) d_tmaps = h->d_tmaps_slots.as<char>() + (size_t)set * h->host_chunk * 128;
    geom.step = (int64_t)rowb;
  }
  SSK_CUDA(cudaEventRecord(h->ev[0], s));

  // ---- registration prep + ECC
  nvtx.next("ssk_stack: registration prep");
  const bool weighted = h->o.accumulation_method == SSK_STACK_WEIGHTED_AVERAGE && h->o.sm_kradius > 0;
  const bool bayer = h->o.accumulation_method == SSK_STACK_BAYER_AVERAGE;
  if (h->o.enable_registration && bayer) {
    // read_input_frame: debayer(raw) keeps the depth, the registration sees the demosaiced frame
    // (c_image_stacking_pipeline_base.cc:221-236, 271-276); the raw samples go to the accumulator below
    const size_t bstep = rowb * 3, bgr_bytes = bstep * h->rows;
    Img g3 = geom;
    g3.cn = 3; g3.step = (int64_t)bstep;
    if (h->reg_h.r.scaled_by_pyrdown() && !h->reg_h.r.flow_enabled() && h->rows >= 4 && h->cols >= 4 && !getenv("SSK_NO_BAYER_PYRDOWN")) {
      // ecc.scale 0.5: the registration only needs pyrDown(gray(debayer(raw))): one pass from the raw samples, the BGR frame
      // is never formed (bit-identical to the chain below)
      if (int e = h->reg_h.r.reserve_batch(n)) return e;
      if (int e = launch_bayer_gray_pyrdown(d_frame_ptrs, h->frames_aligned, geom.step, d, h->rows, h->cols, h->o.bayer_colorid, geom.scale,
                                            h->reg_h.r.ecch.level0_scratch_ptrs(), n, s)) return e;
      h->reg_h.r.ecc_images_ready = true;
      if (int e = h->reg_h.r.prepare(g3, nullptr, n)) return e;
    } else {
      for (int i = 0; i < n; ++i) {
        const void *src = set >= 0 ? (const void *)(h->frame_slots.as<char>() + rowb * h->rows * ((size_t)set * h->host_chunk + i)) : frames[i].data;
        const int64_t sstep = set >= 0 ? (int64_t)rowb : frames[i].step;
        if (int e = launch_debayer_nn2(src, sstep, d, h->rows, h->cols, h->o.bayer_colorid, h->bgr_slots.as<char>() + bgr_bytes * i,
                                       (int64_t)bstep, s)) return e;
      }
      if (int e = h->reg_h.r.prepare(g3, h->d_bgr_ptrs.as<const void *>(), n)) return e;
    }
  } else if (h->o.enable_registration) {
    if (int e = h->reg_h.r.prepare(geom, d_frame_ptrs, n)) return e;
  }
  SSK_CUDA(cudaEventRecord(h->ev[1], s));

  // ---- W1 weights on the unaligned frame
  nvtx.next("ssk_stack: sharpness weights");
  if (weighted) {
    const float *const *Mptrs = nullptr;
    if (h->o.sm_dscale > 0 && std::min(h->rows, h->cols) >= 4) {
      Img cur = geom;
      const void *const *src_ptrs = d_frame_ptrs;
      float *const *dst_ptrs = h->d_half_ptrs.as<float *>();
      int l0 = 0;
      if (h->o.enable_registration && h->reg_h.r.scaled_by_pyrdown() && !h->reg_h.r.normalize_enabled()) {
        // the first cv::pyrDown of the gray frame is exactly the ECC image scaleImage() just produced
        // (c_frame_registration.cc:236-237 vs c_local_variance_sharpness_measure.cc:36): reuse it
        const int nr = (cur.rows + 1) / 2, nc = (cur.cols + 1) / 2;
        float *const *e0 = h->reg_h.r.ecch.level0_scratch_ptrs();
        cur.step = (int64_t)nc * 4; cur.rows = nr; cur.cols = nc; cur.depth = SSK_32F; cur.cn = 1; cur.scale = 1.f;
        Mptrs = e0;
        src_ptrs = reinterpret_cast<const void *const *>(e0);
        l0 = (std::min(nr, nc) < 4) ? h->o.sm_dscale : 1;
      }
      for (int l = l0; l < h->o.sm_dscale; ++l) {
        const int nr = (cur.rows + 1) / 2, nc = (cur.cols + 1) / 2;
        PyrDownArgs pd = {};
        pd.src = cur; pd.src_ptrs = src_ptrs; pd.dst_ptrs = dst_ptrs; pd.dst_rows = nr; pd.dst_cols = nc; pd.batch = n; pd.post_scale = 1.f;
        if (int e = launch_pyrdown(pd, s)) return e;
        cur.step = (int64_t)nc * 4; cur.rows = nr; cur.cols = nc; cur.depth = SSK_32F; cur.cn = 1; cur.scale = 1.f;
        Mptrs = dst_ptrs;
        src_ptrs = reinterpret_cast<const void *const *>(dst_ptrs);
        dst_ptrs = dst_ptrs == h->d_half_ptrs.as<float *>() ? h->d_half2_ptrs.as<float *>() : h->d_half_ptrs.as<float *>();
        if (std::min(nr, nc) < 4) break;
      }
    } else {
      if (int e = launch_to_gray(geom, d_frame_ptrs, nullptr, h->d_half_ptrs.as<float *>(), n, s)) return e;
      Mptrs = h->d_half_ptrs.as<float *>();
    }
    W1Args w = {};
    w.M_ptrs = Mptrs; w.rows = h->w1_rows; w.cols = h->w1_cols; w.kradius = std::max(1, h->o.sm_kradius);
    w.depth_scale = 20.0;   // frames are CV_32F when they reach compute_weights
    w.gmap_ptrs = h->d_gmap_ptrs.as<float *>(); w.partials = h->partials.as<double>(); w.stats = h->stats.as<double>();
    w.out_ptrs = weight_ptrs; w.full_rows = h->rows; w.full_cols = h->cols; w.batch = n;
    w.axis_tab = h->axis_tab.as<int2>(); w.axis_tab_built = &h->axis_tab_built;
    w.uscale = h->o.sm_uscale; w.gmap2_ptrs = h->o.sm_uscale > 0 ? h->d_gmap2_ptrs.as<float *>() : nullptr;
    if (int e = launch_w1(w, s)) return e;
  }
  SSK_CUDA(cudaEventRecord(h->ev[2], s));

  nvtx.next("ssk_stack: register_frame (ECC)");
  if (h->o.enable_registration) {
    if (int e = h->reg_h.r.register_batch(n)) return e;
  }
  SSK_CUDA(cudaEventRecord(h->ev[3], s));

  // ---- fused warp + mask + weights + accumulate
  nvtx.next("ssk_stack: warp + accumulate");
  k_fill_jobs<<<div_up(n, 128), 128, 0, s>>>(h->o.enable_registration ? h->reg_h.r.ecch.device_frames() : nullptr, d_frame_ptrs,
                                             weighted ? weight_ptrs : nullptr,
                                             weighted ? h->stats.as<double>() : nullptr, jobs, n,
                                             h->counter.as<int>(), h->o.enable_registration ? 1 : 0);
  SSK_LAUNCH_CHECK();
  WarpAccArgs a = {};
  a.jobs = jobs; a.njobs = n;
  a.rows = h->out_rows; a.cols = h->out_cols; a.src_rows = h->rows; a.src_cols = h->cols;
  a.src_step = geom.step; a.w_step = (int64_t)h->cols * 4;
  a.depth = d; a.cn = cn; a.scale = geom.scale;
  const ssk_registration_options &ro = h->o.registration;
  a.interp = h->o.enable_registration ? remap_interp(ro.interpolation) : SSK_INTER_NEAREST;
  a.border = h->o.generating_master_frame ? SSK_BORDER_REFLECT101 : ro.border_mode;   // c_image_stacking_pipeline.cc:1647-1650
  for (int i = 0; i < 4; ++i) a.bval[i] = (float)ro.border_value[i];
  a.use_weights = weighted ? 1 : 0;
  a.stage_aligned = h->frames_aligned ? 1 : 0;
  a.side_stream = h->side; a.ev_fork = h->ev_fork; a.ev_join = ev_join; a.defer_join = defer ? 1 : 0;
  if (getenv("SSK_NO_SIDE_STREAM")) a.side_stream = nullptr;   // tuning knob: border-ring kernel in stream order
  {
    ssk_transform t0;
    make_transform(&t0, h->o.enable_registration ? ro.motion_type : SSK_MOTION_TRANSLATION);
    a.map_type = make_mapcoef(t0).type;
  }
  a.tmap_frames = (d_tmaps && !bayer) ? d_tmaps : nullptr;
  a.tmap_weights = (weighted && d_tmaps) ? (par ? h->d_tmaps_weights_b.p : h->d_tmaps_weights.p) : nullptr;
  a.acc = h->acc_h.a.acc.as<float>(); a.wacc = h->acc_h.a.wacc.as<float>();
  if (getenv("SSK_NO_TMA_KERNEL")) a.tmap_frames = a.tmap_weights = nullptr;   // tuning knob: cp.async kernel pair
  if (fused_tma_applicable(a)) a.side_stream = nullptr;                        // one launch over all tiles: nothing to fork
  if (h->o.enable_registration && !bayer && (h->reg_h.r.flow_enabled() || h->upscale != SSK_UPSCALE_NONE)) {
    // per-pixel maps: _current_remap = c_eccflow's refinement of the ECC map (c_frame_registration.cc:900-917) and / or its
    // up-scaling (upscale_remap, c_image_stacking_pipeline.cc:1633-1642)
    if (h->reg_h.r.flow_enabled()) { a.flow = h->reg_h.r.flowh->uv(0); a.flow_stride = (int64_t)h->rows * h->cols; }
    a.side_stream = nullptr;
    if (int e = launch_warp_accumulate_flow(a, h->tab, h->upscale, s)) return e;
  } else if (bayer) {
    // the mask of custom_remap(current_remap, frame, mask, registration_options.interpolation) gates the gather of the raw
    // samples through current_remap (c_image_stacking_pipeline.cc:1644-1651, 1730-1752)
    a.side_stream = nullptr;
    if (h->o.enable_registration) {
      // with enable_eccflow_registration current_remap is the refined per-pixel map (c_frame_registration.cc:900-917)
      if (h->reg_h.r.flow_enabled()) { a.flow = h->reg_h.r.flowh->uv(0); a.flow_stride = (int64_t)h->rows * h->cols; }
      if (int e = launch_bayer_warp_accumulate(a, h->tab, h->o.bayer_colorid, s)) return e;
    } else {
      // no registration: empty remap, no mask -> acc[cc] += src, cntr[cc] += 1 per pixel (c_frame_accumulation.cc:998-1010)
      for (int i = 0; i < n; ++i) {
        BayerAccArgs b = {};
        b.src = geom;
        b.src.data = set >= 0 ? (const void *)(h->frame_slots.as<char>() + rowb * h->rows * ((size_t)set * h->host_chunk + i)) : frames[i].data;
        b.src.step = set >= 0 ? (int64_t)rowb : frames[i].step;
        b.have_map = 0; b.weights = nullptr; b.wtype = -1; b.colorid = h->o.bayer_colorid;
        b.acc = a.acc; b.cntr = a.wacc;
        if (int e = launch_bayer_add(b, s)) return e;
      }
    }
  } else if (int e = launch_warp_accumulate(a, h->tab, s)) return e;
  if (a.side_stream) {
    h->join_recorded[par] = true;
    h->ring_pending = defer ? par : -1;
    SSK_CUDA(cudaEventRecord(h->ev_ring_end, h->side));
  }
  SSK_CUDA(cudaEventRecord(h->ev[4], s));
  if (set >= 0) SSK_CUDA(cudaEventRecord(h->set_free[set], s));
  if (h->o.enable_registration)
    SSK_CUDA(cudaMemcpyAsync(h->rec_all.as<EccFrame>() + rec_off, h->reg_h.r.ecch.device_frames(), sizeof(EccFrame) * n,
                             cudaMemcpyDeviceToDevice, s));
  return SSK_OK;
}

// One chunk of <= max_batch frames: enqueue everything, snapshot the registration records into the chunk's ring slot.
static int stack_submit_chunk(ssk_stack *h, const ssk_mat *frames, int m, int64_t *ticket, bool last_in_call) {
  const int64_t t = h->next_ticket++;
  const int slot = (int)(t % ssk_stack::kRecRing);
  const int base = slot * h->max_batch;
  SSK_CUDA(cudaEventSynchronize(h->rec_ev[slot]));   // the slot's previous download has landed (no-op if never used)
  h->rec_ring_valid[slot] = false;
  if (frames[0].mem == SSK_MEM_DEVICE) {
    if (int e = stack_process_chunk(h, frames, m, -1, base, last_in_call)) return e;
    if (h->ring_pending >= 0) {      // the chunk's ring kernel is still in flight on the side stream
      SSK_CUDA(cudaEventRecord(h->rec_ring_ev[slot], h->side));
      h->rec_ring_valid[slot] = true;
    }
  } else {
    // host frames: upload sub-chunk k+1 on the copy stream while sub-chunk k is processed
    const int hc = h->host_chunk;
    int set = h->set_pos;
    if (int e = stack_upload(h, frames, std::min(hc, m), set)) return e;
    for (int k0 = 0; k0 < m; k0 += hc) {
      const int mk = std::min(hc, m - k0);
      const int next_set = (set + 1) % h->nsets;
      if (k0 + hc < m) {
        if (int e = stack_upload(h, frames + k0 + hc, std::min(hc, m - k0 - hc), next_set)) return e;
      }
      if (int e = stack_process_chunk(h, frames + k0, mk, set, base + k0, true)) return e;
      set = next_set;
    }
    h->set_pos = set;
  }
  if (h->o.enable_registration)
    SSK_CUDA(cudaMemcpyAsync(h->h_rec_all.as<EccFrame>() + base, h->rec_all.as<EccFrame>() + base, sizeof(EccFrame) * m,
                             cudaMemcpyDeviceToHost, h->stream));
  SSK_CUDA(cudaEventRecord(h->rec_ev[slot], h->stream));
  h->rec_n[slot] = m; h->rec_ticket[slot] = t;
  if (ticket) *ticket = t;
  return SSK_OK;
}

static int stack_check_frames(ssk_stack *h, const ssk_mat *frames, int n, int bpp) {
  SSK_REQUIRE(h && frames && n >= 0, "ssk_stack_add_frames: bad argument");
  SSK_REQUIRE(h->have_reference, "ssk_stack: set_reference must be called first");
  SSK_REQUIRE(bpp == h->bpp, "ssk_stack: bpp differs from the reference frame's");
  if (n > 0 && frames[0].type != h->type) {
    // frames of another depth than the reference frame (e.g. a CV_32F master frame over 16-bit input frames): the
    // per-frame buffers follow the frames; the channel count stays the reference's (one for raw Bayer frames)
    SSK_REQUIRE(depth_bytes(type_depth(frames[0].type)) && type_cn(frames[0].type) == type_cn(h->type),
                "ssk_stack: frame type differs from the reference frame's channel layout");
    if (int e = ssk_stack_sync(h)) return e;
    h->type = frames[0].type;
    if (int e = stack_alloc_slots(h)) return e;
  }
  for (int i = 0; i < n; ++i) {
    if (int e = stack_check_mat(&frames[i], "ssk_stack: frame")) return e;
    SSK_REQUIRE(frames[i].rows == h->rows && frames[i].cols == h->cols && frames[i].type == h->type,
                "ssk_stack: frame geometry/type differs from the reference frame");
  }
  return SSK_OK;
}

int ssk_stack_submit(ssk_stack *h, const ssk_mat *frames, int n, int bpp, int64_t *ticket) {
  if (int e = stack_check_frames(h, frames, n, bpp)) return e;
  SSK_REQUIRE(n >= 1 && n <= h->max_batch, "ssk_stack_submit: 1..max_batch frames per call");
  return stack_submit_chunk(h, frames, n, ticket, true);
}

int ssk_stack_wait(ssk_stack *h, int64_t ticket, ssk_transform *transforms_out, ssk_ecc_status *status_out, int capacity,
                   int *n_out) {
  SSK_REQUIRE(h, "null handle");
  const int slot = (int)(ticket % ssk_stack::kRecRing);
  SSK_REQUIRE(ticket >= 0 && h->rec_ticket[slot] == ticket, "ssk_stack_wait: unknown ticket (only the last 4 chunks are kept)");
  SSK_CUDA(cudaEventSynchronize(h->rec_ev[slot]));
  if (h->rec_ring_valid[slot]) SSK_CUDA(cudaEventSynchronize(h->rec_ring_ev[slot]));   // caller's device frames are free again
  const int m = h->rec_n[slot];
  if (n_out) *n_out = m;
  if (h->o.enable_registration && (transforms_out || status_out)) {
    SSK_REQUIRE(capacity >= m, "ssk_stack_wait: result arrays too small");
    const EccFrame *f = h->h_rec_all.as<EccFrame>() + (size_t)slot * h->max_batch;
    for (int i = 0; i < m; ++i) {
      if (transforms_out) transforms_out[i] = f[i].t;
      if (status_out) {
        ssk_ecc_status &st = status_out[i];
        st.rho = f[i].rho; st.min_rho = h->o.registration.ecc.min_rho; st.eps = f[i].eps;
        st.num_iterations = f[i].num_iterations; st.max_iterations = h->o.registration.ecc.max_iterations;
        st.ok = f[i].ok; st.failed = f[i].failed;
      }
    }
  }
  return SSK_OK;
}

int ssk_stack_add_frames_async(ssk_stack *h, const ssk_mat *frames, int n, int bpp) {
  if (int e = stack_check_frames(h, frames, n, bpp)) return e;
  for (int i0 = 0; i0 < n; i0 += h->max_batch) {
    if (int e = stack_submit_chunk(h, frames + i0, std::min(h->max_batch, n - i0), nullptr, i0 + h->max_batch >= n)) return e;
  }
  return SSK_OK;
}

int ssk_stack_flush(ssk_stack *h) {
  SSK_REQUIRE(h, "null handle");
  if (h->ring_pending >= 0) {
    SSK_CUDA(cudaStreamWaitEvent(h->stream, h->ring_pending ? h->ev_join_b : h->ev_join, 0));
    h->ring_pending = -1;
  }
  return SSK_OK;
}

int ssk_stack_sync(ssk_stack *h) {
  SSK_REQUIRE(h, "null handle");
  if (int e = ssk_stack_flush(h)) return e;
  SSK_CUDA(cudaMemcpyAsync(h->h_counter.p, h->counter.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  SSK_CUDA(cudaStreamSynchronize(h->stream));
  if (h->side) SSK_CUDA(cudaStreamSynchronize(h->side));   // a ring kernel left behind by a call that failed half-way
  h->acc_h.a.frames = *h->h_counter.as<int>();
  return SSK_OK;
}

int ssk_stack_add_frames(ssk_stack *h, const ssk_mat *frames, int n, int bpp, ssk_transform *transforms_out,
                         ssk_ecc_status *status_out) {
  if (int e = stack_check_frames(h, frames, n, bpp)) return e;
  for (int i0 = 0; i0 < n; i0 += h->max_batch) {
    const int m = std::min(h->max_batch, n - i0);
    int64_t t = -1;
    if (int e = stack_submit_chunk(h, frames + i0, m, &t, i0 + h->max_batch >= n)) return e;
    if (int e = ssk_stack_wait(h, t, transforms_out ? transforms_out + i0 : nullptr, status_out ? status_out + i0 : nullptr, m, nullptr)) return e;
  }
  return ssk_stack_sync(h);
}

// Start of a new run over the same reference: the pipeline creates a fresh accumulator (create_frame_accumulation,
// c_image_stacking_pipeline.cc:450-466); here the device accumulator and the frame counter are zeroed in stream order.
int ssk_stack_reset(ssk_stack *h) {
  SSK_REQUIRE(h && h->have_reference, "ssk_stack_reset: set_reference must be called first");
  if (int e = ssk_stack_flush(h)) return e;
  Acc &a = h->acc_h.a;
  const size_t npix = (size_t)a.rows * a.cols;
  const bool bayer = a.kind == SSK_ACC_BAYER_AVERAGE;
  SSK_CUDA(cudaMemsetAsync(a.acc.p, 0, npix * (bayer ? 3 : a.cn) * 4, h->stream));
  SSK_CUDA(cudaMemsetAsync(a.wacc.p, 0, npix * (bayer ? 3 : 1) * 4, h->stream));
  SSK_CUDA(cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->stream));
  a.frames = 0;
  return SSK_OK;
}

int ssk_stack_compute(ssk_stack *h, ssk_mat *avg, ssk_mat *mask) {
  SSK_REQUIRE(h, "null handle");
  if (int e = ssk_stack_sync(h)) return e;
  return ssk_acc_compute(&h->acc_h, avg, mask, 1.0);
}

int ssk_stack_compute_inpainted(ssk_stack *h, ssk_mat *avg, ssk_mat *mask, int max_levels) {
  SSK_REQUIRE(h, "null handle");
  if (int e = ssk_stack_sync(h)) return e;
  return ssk_acc_compute_inpainted(&h->acc_h, avg, mask, 1.0, max_levels);
}

int ssk_stack_accumulated_frames(ssk_stack *h) {
  if (!h) return 0;
  if (ssk_stack_sync(h)) return -1;
  return h->acc_h.a.frames;
}

ssk_acc *ssk_stack_accumulator(ssk_stack *h) {
  if (!h) return nullptr;
  ssk_stack_flush(h);     // accumulator calls are ordered on the handle's stream: the pending ring kernel first
  return &h->acc_h;
}
ssk_reg *ssk_stack_registration(ssk_stack *h) { return h ? &h->reg_h : nullptr; }

void *ssk_stack_stream(ssk_stack *h) { return h ? (void *)h->stream : nullptr; }

int ssk_stack_stage_times(ssk_stack *h, float ms[4]) {
  SSK_REQUIRE(h && ms, "null argument");
  if (int e = ssk_stack_sync(h)) return e;
  float t[4];
  for (int i = 0; i < 4; ++i) SSK_CUDA(cudaEventElapsedTime(&t[i], h->ev[i], h->ev[i + 1]));
  // the warp+accumulate stage ends when both of its kernels have: the ring kernel runs on the side stream
  if (h->side && h->join_recorded[0]) {
    float tr = 0.f;
    if (cudaEventElapsedTime(&tr, h->ev[3], h->ev_ring_end) == cudaSuccess) t[3] = std::max(t[3], tr);
    else cudaGetLastError();
  }
  ms[0] = t[0]; ms[1] = t[1]; ms[2] = t[2]; ms[3] = t[3];
  return SSK_OK;
}

}  // extern "C"
