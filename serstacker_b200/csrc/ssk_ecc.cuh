// Internal interface of the persistent ECC registration kernel (ssk_ecc.cu).
#pragma once
#include "ssk_common.cuh"

namespace ssk {

constexpr int kMaxLevels = 12;
constexpr int kTmaLevels = 2;   // finest levels that may take the TMA-staged inverse-compositional pass
constexpr int kTraceRec = 40;   // level, pass, num_it, err, newerr, lambda, eps, n | t[8] | tq[8] | deltap[8] | v[8]

// One pyramid level, reference side (shared by every frame of every batch).
struct EccLevel {
  int cols, rows;
  const float *ref;         // smoothed reference image of this level (dense)
  const uint8_t *refmask;   // reference mask of this level as the solver keeps it (eroded where the solver erodes) or null
  const float *gx, *gy;     // masked reference gradients (INVERSE_COMPOSITIONAL*), dense
  int64_t cur_off;          // offset (floats) of this level inside a frame's pyramid buffer
  double RMA;               // reference mask area
};

// Cached reference-side normal matrices of the inverse-compositional solvers.
struct EccHpCache {
  float Hp[kMaxLevels][64];     // M x M row-major
  float jp[kMaxLevels][12];     // parameters (8) + aux (4) the steepest-descent images were evaluated with
  int valid[kMaxLevels];
};

struct alignas(64) EccConfig {
  // TMA-staged passes (ssk_ecc_impl.cuh::pass_ic_tma): tensor maps of the tma_levels finest levels - [0] the current images of
  // all frame slots (3-D: x, y, slot; slot i = pyr_base + i * pyr_floats), [1] reference, [2] gx, [3] gy (2-D).  0 = off.
  alignas(64) unsigned char tm[kTmaLevels][4][128];
  int tma_levels;
  const float *pyr_base;
  int64_t pyr_floats;
  int nlevels;
  EccLevel lv[kMaxLevels];
  int method;                 // SSK_ECC_*
  int interp;                 // interpolation of the FORWARD_ADDITIVE image warp (c_ecc_align::_interpolation)
  int max_iterations;
  double update_step_scale;
  double epsx;                // level-0 stop threshold; doubled per coarser level (ecc2.cc:1030-1034)
  double max_epse;            // c_ecc_align::_max_epse = 1e-4
  // c_frame_registration flow (c_frame_registration.cc:797-872); all zero for a bare c_ecch::align
  int motion_type;            // type of the transform being estimated
  int translation_first;      // ecch_estimate_translation_first && motion != TRANSLATION && ecch_max_level != 0
  int check_rho;              // compute_correlation gate
  double min_rho;
  double final_scale;         // scale_transfrom(1/ecc.scale) applied on success (1 = none)
  // inverse-compositional caches (device pointers)
  const EccHpCache *hp_trans; // translation-first pass (M = 2)
  EccHpCache *hp_main;        // main transform
  int hp_main_mode;           // 0: use cache; 1: recompute per frame at level start; 2: recompute and store (first frame)
  // optional per-trial trace of frame 0 (debug aid, like the reference's debug dumps): kTraceRec floats per record
  float *trace; int trace_capacity; int *trace_count;
};

// Per-frame input/output record (device memory).
struct EccFrame {
  const float *pyr;           // this frame's current-image pyramid (level l at pyr + lv[l].cur_off)
  ssk_transform t;            // in: start transform; out: estimated transform
  double rho, eps;
  int num_iterations;
  int ok;                     // 1 registered, 0 dropped
  int failed;                 // solver failure flag
  int pad;
  MapCoef map;                // out: full-resolution map coefficients of `t` (for the fused warp kernel)
  const uint8_t *cmask;       // this frame's current-mask pyramid as its solver keeps it (level l at cmask + lv[l].cur_off
                              // bytes), or null: no current mask (c_ecc_align::set_current_image, ecc2.cc:611-632, 1214-1234)
};

// box geometry of the TMA-staged passes (defined next to the kernel, ssk_ecc.cu)
int ecc_tma_tile_w(); int ecc_tma_tile_h(); int ecc_tma_win_w(); int ecc_tma_win_h();

// Launches one thread-block cluster per frame; the cluster runs the whole coarse-to-fine alignment of its
// frame (all levels, all iterations, the correlation gate) on the device.
int launch_ecc(const EccConfig &cfg, EccFrame *frames, int nframes, int cluster_size, cudaStream_t s);

// Initialises `n` frame records on the device: t = t0, pyr = pyr_base + i * pyr_floats, status cleared.
int launch_ecc_init_frames(EccFrame *frames, int n, const ssk_transform &t0, const float *pyr_base, int64_t pyr_floats,
                           const uint8_t *mask_base, cudaStream_t s);

// Reference-side precompute: Hp of every level for the translation transform and, when the steepest-descent
// images do not depend on the parameters (translation / affine), for the main transform.
int launch_ecc_precompute(const EccConfig &cfg, EccHpCache *hp_trans, EccHpCache *hp_main, cudaStream_t s);

}  // namespace ssk
