// c_eccflow on the device (ssk_eccflow.cu): dense smooth optical flow that refines a registration map per pixel.
// Reference: core/proc/image_registration/ecc2.h:515-662, ecc2.cc:2220-2865; used by c_frame_registration::register_frame
// (c_frame_registration.cc:900-917) when enable_eccflow_registration is set.
#pragma once
#include <vector>
#include "ssk_engine.cuh"

namespace ssk {

constexpr int kMaxFlowLevels = 64;

// cv::resize(INTER_AREA) tables of one axis (computeResizeAreaTab): dst index d gathers count[d] source samples from start[d]
// with the weights alpha[off[d] ...]
struct FlowAreaAxis { const int *start, *count, *off; const float *alpha; };
// cv::resize(INTER_CUBIC) tables of one axis: first tap = s[d] - 1, Keys coefficients c[d] (interpolateCubic, A = -0.75)
struct FlowCubicAxis { const int *s; const float4 *c; };

class EccFlow {
 public:
  ssk_eccflow_options opts;
  cudaStream_t stream = nullptr;
  int nlevels = 0;
  int lw[kMaxFlowLevels], lh[kMaxFlowLevels], lsrc[kMaxFlowLevels];   // level size, level it is reduced from
  int cw[kMaxFlowLevels], ch[kMaxFlowLevels];                          // size of the avgdown() grid of the level
  int64_t loff[kMaxFlowLevels], coff[kMaxFlowLevels];                  // element offsets of a level in a pyramid / coarse buffer
  int64_t pyr_px = 0, coarse_px = 0;
  bool have_reference = false, have_ref_mask = false, cur_mask_set = false;
  int capacity = 0;

  int init(const ssk_eccflow_options &o, cudaStream_t s);
  // reference: dense CV_32FC1 device image; d_mask: dense CV_8UC1 device mask of the same size or null
  int set_reference(const float *d_img, int rows, int cols, const uint8_t *d_mask);
  int reserve(int batch);
  // level 0 of the current image of slot b (dense CV_32FC1, written by the caller before build_current)
  float *level0(int b) { return cur_pyr.as<float>() + (int64_t)b * pyr_px; }
  float *const *level0_ptrs() { return d_cur_ptrs.as<float *>(); }        // device array [capacity]
  // current pyramids of `batch` frames from their level-0 images; d_mask: dense level-0 mask of frame 0 (batch == 1) or null
  int build_current(int batch, const uint8_t *d_mask);
  // compute_uv: frames[b].map (device records of the ECC kernel) or rmap0 (explicit CV_32FC2 map of frame 0, dense) or
  // neither (zero initial flow); the flow of frame b ends in uv(b) (level-0 size, dense CV_32FC2)
  int compute(int batch, const EccFrame *d_frames, const float2 *d_rmap0);
  float2 *uv(int b) { return uv_a.as<float2>() + (int64_t)b * lw[0] * lh[0]; }
  // rmap = uv + (x, y) (ecc_flow_to_remap) for frame b into a dense device buffer
  int write_remap(int b, float2 *d_rmap);
  // debug / test access: which = 0 reference, 1 current (slot 0), 2 Ix, 3 Iy (level size, CV_32FC1); 4 = D (coarse, CV_32FC4)
  const float *image(int which, int level) const;

 private:
  DevBuf ref_pyr, ref_ix, ref_iy, ref_D, ref_mask, cur_pyr, cur_mask, uv_a, uv_b, raw, cuv, tabs, d_cur_ptrs, d_lvl_ptrs;
  // per level: tables live in `tabs`
  struct LevelTabs {
    FlowAreaAxis ax, ay;            // level -> coarse grid
    FlowCubicAxis ux, uy;           // coarse grid -> level
    FlowCubicAxis nx, ny;           // next coarser level -> level
  };
  LevelTabs lt[kMaxFlowLevels];
  FlowCubicAxis ix, iy;             // level 0 -> last level (initial flow)
  int build_tables();
  int reduce_smem_optin = 0;
};

}  // namespace ssk

struct ssk_eccflow {
  ssk::EccFlow f;
  cudaStream_t stream = nullptr;
  ssk::DevBuf st_img, st_mask, st_map;
  ~ssk_eccflow() { if (stream) cudaStreamDestroy(stream); }
};
