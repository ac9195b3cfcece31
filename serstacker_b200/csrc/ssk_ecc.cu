// Dispatch of the persistent ECC registration kernel + reference-side precompute (see ssk_ecc_impl.cuh).
#include "ssk_ecc_impl.cuh"

namespace ssk {

int launch_ecc_fa(const EccConfig &, EccFrame *, int, int, cudaStream_t);
int launch_ecc_ic(const EccConfig &, EccFrame *, int, int, cudaStream_t);
int launch_ecc_lm(const EccConfig &, EccFrame *, int, int, cudaStream_t);
int launch_ecc_iclm(const EccConfig &, EccFrame *, int, int, cudaStream_t);

namespace {

// Hp of every level for translation and (when parameter independent) for the main transform
template <int TYPE>
__device__ void precompute_type(Ctx &c, EccHpCache *out) {
  constexpr int M = NParams<TYPE>::M;
  Shared &S = *c.S;
  for (int lvl = 0; lvl < c.cfg->nlevels; ++lvl) {
    pass_hp<TYPE>(c, lvl);
    if (c.tid == 0 && c.rank == 0) {
      float H[64];
      unpack_H(M, S.tot, H);
      for (int i = 0; i < M * M; ++i) out->Hp[lvl][i] = H[i];
      for (int i = 0; i < 12; ++i) out->jp[lvl][i] = 0.f;
      out->valid[lvl] = 1;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NT, 1) k_ecc_precompute(const __grid_constant__ EccConfig cfg, EccHpCache *hp_trans,
                                                         EccHpCache *hp_main) {
  __shared__ Shared S;
  cg::cluster_group cluster = cg::this_cluster();
  Ctx c;
  c.cfg = &cfg; c.csize = (int)cluster.num_blocks(); c.rank = (int)cluster.block_rank(); c.tid = threadIdx.x; c.buf = 0;
  c.S = &S; c.frame = nullptr;
  if (c.tid == 0) for (int i = 0; i < 8; ++i) S.jc.c[i] = 0.f;
  __syncthreads();
  if (hp_trans) precompute_type<SSK_MOTION_TRANSLATION>(c, hp_trans);
  if (hp_main) {
    if (cfg.motion_type == SSK_MOTION_AFFINE) precompute_type<SSK_MOTION_AFFINE>(c, hp_main);
    else if (cfg.motion_type == SSK_MOTION_TRANSLATION) precompute_type<SSK_MOTION_TRANSLATION>(c, hp_main);
  }
  cluster.sync();
}

__global__ void k_ecc_init_frames(EccFrame *frames, int n, const ssk_transform t0, const float *pyr_base, int64_t pyr_floats,
                                  const uint8_t *mask_base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  EccFrame &f = frames[i];
  f.pyr = pyr_base + (int64_t)i * pyr_floats;
  f.t = t0;
  f.rho = -1; f.eps = 0; f.num_iterations = 0; f.ok = 0; f.failed = 0; f.pad = 0;
  f.map = make_mapcoef(t0);
  f.cmask = mask_base ? mask_base + (int64_t)i * pyr_floats : nullptr;
}

}  // namespace

int launch_ecc_init_frames(EccFrame *frames, int n, const ssk_transform &t0, const float *pyr_base, int64_t pyr_floats,
                           const uint8_t *mask_base, cudaStream_t s) {
  k_ecc_init_frames<<<div_up(n, 128), 128, 0, s>>>(frames, n, t0, pyr_base, pyr_floats, mask_base);
  SSK_LAUNCH_CHECK();
  return SSK_OK;
}

int ecc_tma_tile_w() { return TL_W; }
int ecc_tma_tile_h() { return TL_H; }
int ecc_tma_win_w() { return TL_WW; }
int ecc_tma_win_h() { return TL_WH; }

int launch_ecc(const EccConfig &cfg, EccFrame *frames, int nframes, int cluster_size, cudaStream_t s) {
  SSK_REQUIRE(cluster_size >= 1 && cluster_size <= 8, "ECC: cluster size 1..8");
  SSK_REQUIRE(cfg.nlevels >= 1 && cfg.nlevels <= kMaxLevels, "ECC: pyramid depth");
  SSK_REQUIRE(cfg.interp == SSK_INTER_LINEAR || cfg.interp == SSK_INTER_NEAREST,
              "ECC forward-additive warp: LINEAR or NEAREST interpolation");
  switch (cfg.method) {
    case SSK_ECC_FORWARD_ADDITIVE: return launch_ecc_fa(cfg, frames, nframes, cluster_size, s);
    case SSK_ECC_INVERSE_COMPOSITIONAL: return launch_ecc_ic(cfg, frames, nframes, cluster_size, s);
    case SSK_ECC_INVERSE_COMPOSITIONAL_LM: return launch_ecc_iclm(cfg, frames, nframes, cluster_size, s);
    default: return launch_ecc_lm(cfg, frames, nframes, cluster_size, s);
  }
}

int launch_ecc_precompute(const EccConfig &cfg, EccHpCache *hp_trans, EccHpCache *hp_main, cudaStream_t s) {
  void *args[3] = {(void *)&cfg, (void *)&hp_trans, (void *)&hp_main};
  return launch_clustered((const void *)k_ecc_precompute, args, 1, 4, s);
}

}  // namespace ssk
