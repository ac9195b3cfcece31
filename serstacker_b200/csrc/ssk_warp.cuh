// Internal launch interface of the warp / accumulate kernels (ssk_warp.cu).
#pragma once
#include "ssk_common.cuh"

namespace ssk {

// One frame of a batch as the fused warp+accumulate kernel sees it.  The registration kernel writes
// `map` and `ok` on the device, so a batch never round-trips through the host.
struct FrameJob {
  const void *frame;      // device pointer, geometry/type common to the batch
  const float *weights;   // full-resolution CV_32FC1 weight map (or null)
  MapCoef map;            // reference pixel -> frame pixel
  int ok;                 // 0: frame dropped (registration failed)
  int pad;
};

struct WarpAccArgs {
  const FrameJob *jobs;   // device array [njobs]
  int njobs;
  int rows, cols;         // output (= accumulator = reference frame) size
  int src_rows, src_cols; // frame size
  int64_t src_step;       // bytes
  int64_t w_step;         // bytes (weight maps)
  int depth, cn;
  float scale;            // sample scale for integer frames
  int interp;             // SSK_INTER_*
  int border;             // SSK_BORDER_* for the frame
  float bval[4];
  int use_weights;        // 1: weighted_average (w = remap(weights, interp, CONSTANT 0) * mask)
  int stage_aligned;      // all frame / weight-map base pointers are 16-byte aligned (enables cp.async staging)
  int map_type;           // MAP_* common to all jobs of the batch (-1: mixed / unknown -> generic kernel)
  void *side_stream, *ev_fork, *ev_join;   // host only: optional side stream (+2 events) for the border-ring kernel
  int defer_join;         // host only: do not make `stream` wait for the ring kernel (the caller joins ev_join later)
  // optional TMA staging of the interior tiles (32F frames + weight maps): device arrays of 128-byte tensor maps, one per
  // job, box = staged_box_w() x staged_box_h() elements; null -> cp.async staging
  const void *tmap_frames, *tmap_weights;
  float *acc;             // running mean, rows x cols x cn (dense)
  float *wacc;            // running weight sum, rows x cols (dense)
  // per-pixel maps (c_eccflow): the map of job j is flow[j * flow_stride + y * cols + x] + (x, y); null: analytic maps (jobs[j].map)
  const float2 *flow; int64_t flow_stride;
};

int launch_warp_accumulate(const WarpAccArgs &a, const Tables &tab, cudaStream_t stream);
// TMA form (ssk_fused_tma.cu): CV_32F single-channel frames with tensor maps, affine-like maps; one launch over all tiles.
bool fused_tma_applicable(const WarpAccArgs &a);
int launch_warp_accumulate_tma(const WarpAccArgs &a, const Tables &tab, cudaStream_t stream);
// per-pixel-map form (ssk_fused_flow.cu): a.flow holds the c_eccflow result of every job (or null: analytic maps) and / or the
// map is up-scaled (frame_upscale_option, c_image_stacking_pipeline.cc:1633-1660: a.rows x a.cols is the up-scaled size)
int launch_warp_accumulate_flow(const WarpAccArgs &a, const Tables &tab, int upscale_option, cudaStream_t stream);
// Bayer form of the fused loop (ssk_bayer.cu): a.jobs[j].frame are raw Bayer frames, a.acc / a.wacc the rows x cols x 3 sums
// and counters of c_bayer_average; the mask is base_remap's eroded validity for a.interp.
int launch_bayer_warp_accumulate(const WarpAccArgs &a, const Tables &tab, int colorid, cudaStream_t stream);
// geometry of the staged source window of an interior tile (32F frames), for the tensor maps of the TMA path
int staged_box_w();
int staged_box_h();
bool encode_tmap_2d_f32(void *out128, const void *base, int cols, int rows, int64_t step_bytes, int box_w, int box_h);
bool encode_tmap_3d_f32(void *out128, const void *base, int cols, int rows, int depth, int64_t step_bytes, int64_t slice_bytes, int box_w,
                        int box_h);

// cv::remap of a CV_32F image (cn 1..4) by an analytic map or an explicit CV_32FC2 map.
struct RemapArgs {
  Img src;
  float *dst; int64_t dst_step;          // CV_32F, same cn
  int rows, cols;                        // dst size
  MapCoef map; const float2 *rmap; int64_t rmap_step;   // rmap != null overrides map
  int interp, border; float bval[4];
};
int launch_remap(const RemapArgs &a, const Tables &tab, cudaStream_t stream);

// dst_mask = erode5x5(remap(src_mask or all-255, interp, CONSTANT 0) >= 255), border value 255.
struct RemapMaskArgs {
  const uint8_t *src_mask; int64_t src_mask_step;   // null => all-255 source
  int src_rows, src_cols;
  uint8_t *dst; int64_t dst_step;
  uint8_t *tmp;                                    // rows*cols scratch (pre-erode)
  int rows, cols;
  MapCoef map; const float2 *rmap; int64_t rmap_step;
  int interp;
};
int launch_remap_mask(const RemapMaskArgs &a, const Tables &tab, cudaStream_t stream);

// compute_ellipsoid_zrotation_remap (core/proc/feature2d/ellipsoid.cc:206-277): derotation map of a planetary
// ellipsoid from pose R1 (as imaged) to pose R2 (target), disk mask and limb-darkening weight (before its remap).
struct EllipsoidArgs {
  int rows, cols;
  double cx, cy, A, B, C;
  double R1[9], R2[9];            // row-major 3x3, XYZscreen = R * XYZplanet
  int bx, by, bw, bh;             // ellipse crop box (ellipse_crop_box of ellipsoid_bbox under R2)
  double ca, sa;                  // cos / sin of the bounding ellipse angle
  double wscale;
  float2 *rmap; float *wmap; uint8_t *rmask;   // dense outputs
};
int launch_ellipsoid_remap(const EllipsoidArgs &a, cudaStream_t s);

// Per-frame weight of c_jdr_pipeline::derotate_and_average_frames before its smoothing (c_jdr_pipeline.cc:1207-1227):
// w = wmap (0 below 1e-5) [* lpg]; master frame: 1 outside the disk mask; 0 where the frame mask is 0.  In place on w.
int launch_jdr_weights(float *w, const float *lpg_map, const uint8_t *rmask, const uint8_t *mask, int64_t mask_step,
                       int rows, int cols, int is_master, cudaStream_t s);

// the same weight with its two remaps fused (w = remap(wpre, rmap, LINEAR, CONSTANT) [* remap(lpg, rmap, LINEAR, TRANSPARENT)]),
// and the frame's TRANSPARENT derotation fused with the weighted add (dense CV_32FC1 images of rows x cols)
int launch_jdr_weights_fused(const float *wpre, const float *lpg_map, const float2 *rmap, const uint8_t *rmask, const uint8_t *mask,
                             int64_t mask_step, int rows, int cols, int is_master, float *w, cudaStream_t s);
int launch_jdr_remap_add(const float *frame, const float2 *rmap, const float *weights, int rows, int cols, float *acc, float *wacc,
                         cudaStream_t s);

// c_weigthed_average::add without warp (c_frame_accumulation.cc:20-129)
struct AccAddArgs {
  Img src;
  const void *weights; int64_t w_step; int wtype;  // -1 none, SSK_8UC1, SSK_32FC1
  float *acc; float *wacc;
};
int launch_acc_add(const AccAddArgs &a, cudaStream_t stream);

// compute(): avg = acc * dscale, mask = W > 0
int launch_acc_compute(const float *acc, const float *wacc, int rows, int cols, int cn, float dscale,
                       float *avg, int64_t avg_step, uint8_t *mask, int64_t mask_step, cudaStream_t stream);
// running mean <-> sum form
int launch_acc_sum_form(float *acc, const float *wacc, int64_t npix, int cn, int to_sum, cudaStream_t stream);

// c_bayer_average (c_frame_accumulation.cc:988-1126 / 1205-1238)
struct BayerAccArgs {
  Img src;                                   // raw bayer, cn = 1
  MapCoef map; const float2 *rmap; int64_t rmap_step; int have_map;
  const void *weights; int64_t w_step; int wtype;
  int colorid;
  float *acc; float *cntr;                   // rows x cols x 3
};
int launch_bayer_add(const BayerAccArgs &a, cudaStream_t stream);
int launch_bayer_compute(const float *acc, const float *cntr, int rows, int cols, float *avg, int64_t avg_step,
                         uint8_t *mask, int64_t mask_step, cudaStream_t stream);

}  // namespace ssk
