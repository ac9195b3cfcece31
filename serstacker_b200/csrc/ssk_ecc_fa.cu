// ECC registration kernel instantiations for one solver (see ssk_ecc_impl.cuh).
#include "ssk_ecc_impl.cuh"
namespace ssk {
int launch_ecc_fa(const EccConfig &cfg, EccFrame *frames, int nframes, int cluster_size, cudaStream_t s) {
  return launch_ecc_method<SSK_ECC_FORWARD_ADDITIVE>(cfg, frames, nframes, cluster_size, s);
}
}  // namespace ssk
