// Probe of cp.async.bulk.tensor.2d behaviour on this GPU: negative / unaligned / fully out-of-bounds box coordinates.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/probes/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap tm, const void *tmg, int use_global, int x, int y, int buf, float *out) {
  __shared__ __align__(128) float s[2][40 * 40];
  __shared__ __align__(8) unsigned long long mbar[2];
  const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar[buf]);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const void *d = use_global ? tmg : (const void *)&tm;
    if (use_global) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(d) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(6400u) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(s[buf])),
                 "l"(d), "r"(x), "r"(y), "r"(mb)
                 : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mb) : "memory");
  for (int i = threadIdx.x; i < 1600; i += blockDim.x) out[i] = s[buf][i];
}

int main() {
  const int W = 64, H = 48;
  std::vector<float> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = 1.0f + i;
  float *d, *o;
  cudaMalloc(&d, W * H * 4);
  cudaMalloc(&o, 1600 * 4);
  cudaMemcpy(d, h.data(), W * H * 4, cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeFn fn = (EncodeFn)sym;
  CUtensorMap tm;
  const cuuint64_t dims[2] = {W, H}, strides[1] = {W * 4};
  const cuuint32_t box[2] = {40, 40}, es[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  void *tmg;
  cudaMalloc(&tmg, 128);
  cudaMemcpy(tmg, &tm, 128, cudaMemcpyHostToDevice);
  const int cases[][2] = {{0, 0}, {4, 3}, {-4, 0}, {0, -1}, {-8, -5}, {28, 21}, {60, 45}, {64, 48}, {100, 101}, {-40, -40}, {-44, 0}, {3, 0}};
  for (int g = 0; g < 2; ++g)
    for (int buf = 0; buf < 2; ++buf)
      for (auto &c : cases) {
        k<<<1, 128>>>(tm, tmg, g, c[0], c[1], buf, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("global=%d buf=%d (%d,%d): %s\n", g, buf, c[0], c[1], cudaGetErrorString(e)); return 1; }
        std::vector<float> out(1600);
        cudaMemcpy(out.data(), o, 6400, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < 40; ++yy)
          for (int xx = 0; xx < 40; ++xx) {
            const int gx = c[0] + xx, gy = c[1] + yy;
            const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * W + gx] : 0.f;
            if (out[yy * 40 + xx] != want) ++bad;
          }
        printf("global=%d buf=%d (%d,%d): ok, mismatches %d\n", g, buf, c[0], c[1], bad);
      }
  return 0;
}
