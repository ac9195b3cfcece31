// ECC registration kernel instantiations for one solver (see ssk_ecc_impl.cuh).
#include "ssk_ecc_impl.cuh"
namespace ssk {
int launch_ecc_lm(const EccConfig &cfg, EccFrame *frames, int nframes, int cluster_size, cudaStream_t s) {
  return launch_ecc_method<SSK_ECC_LM>(cfg, frames, nframes, cluster_size, s);
}
}  // namespace ssk
