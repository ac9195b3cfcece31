// Process-wide runtime state: error string, launch counter, remap tables.
#include "ssk_common.cuh"
#include <cuda.h>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <mutex>
#include <vector>

namespace ssk {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string &msg) { g_err = msg; }
const std::string &last_error() { return g_err; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_err = buf;
  return SSK_ERR_CUDA;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only).
// A 2-D fp32 image (cols x rows, row pitch step_bytes) with a box of box_w x box_h elements, no swizzle, out-of-bounds
// elements read as zero.  Returns false when the driver entry point is missing or the geometry is not encodable
// (base not 16-byte aligned, pitch not a multiple of 16); the caller then stays on the cp.async staging path.
bool encode_tmap_2d_f32(void *out128, const void *base, int cols, int rows, int64_t step_bytes, int box_w, int box_h) {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
    cudaGetLastError();
  }
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (step_bytes & 15) || box_w > 256 || box_h > 256 || ((box_w * 4) & 15)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)step_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estr[2] = {1, 1};
  return fn(static_cast<CUtensorMap *>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The same for a stack of `depth` images of one geometry, slice_bytes apart (x, y, slice): box = box_w x box_h x 1.
bool encode_tmap_3d_f32(void *out128, const void *base, int cols, int rows, int depth, int64_t step_bytes, int64_t slice_bytes, int box_w,
                        int box_h) {
  unsigned char probe[128];
  if (!encode_tmap_2d_f32(probe, base, cols, rows, step_bytes, box_w, box_h)) return false;    // resolves the entry point, checks the 2-D part
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  if ((slice_bytes & 15) || depth < 1) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)depth};
  const cuuint64_t strides[2] = {(cuuint64_t)step_bytes, (cuuint64_t)slice_bytes};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return reinterpret_cast<EncodeFn>(sym)(static_cast<CUtensorMap *>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims,
                                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- stream-ordered call chains (ssk_set_stream_ordered, include/ssk.h) ----------------------------------------------
// Handles and the stateless operators own different streams.  In stream-ordered mode a call whose matrices all live on the
// device ends with an event record instead of a host wait; the next such call makes its stream wait for that event, so
// the library's calls stay totally ordered on the device while the host runs ahead.
namespace {
std::atomic<int> g_ordered{0};
struct Chain { cudaEvent_t ev = nullptr, hop = nullptr; bool pending = false; };
thread_local Chain g_chain;
}  // namespace

int stream_ordered() { return g_ordered.load(std::memory_order_relaxed); }
int set_stream_ordered(int enable) { return g_ordered.exchange(enable ? 1 : 0); }

int chain_wait(cudaStream_t s) {
  if (g_chain.pending) SSK_CUDA(cudaStreamWaitEvent(s, g_chain.ev, 0));
  return SSK_OK;
}

int chain_finish(cudaStream_t s, bool all_device) {
  if (stream_ordered() && all_device) {
    if (!g_chain.ev) SSK_CUDA(cudaEventCreateWithFlags(&g_chain.ev, cudaEventDisableTiming));
    SSK_CUDA(cudaEventRecord(g_chain.ev, s));
    g_chain.pending = true;
    return SSK_OK;
  }
  SSK_CUDA(cudaStreamSynchronize(s));
  return SSK_OK;
}

int chain_drain() {
  if (g_chain.pending) {
    SSK_CUDA(cudaEventSynchronize(g_chain.ev));
    g_chain.pending = false;
  }
  return SSK_OK;
}

// `waiter` continues after everything `producer` has been given so far (no host wait)
int stream_after(cudaStream_t waiter, cudaStream_t producer) {
  if (waiter == producer) return SSK_OK;
  if (!g_chain.hop) SSK_CUDA(cudaEventCreateWithFlags(&g_chain.hop, cudaEventDisableTiming));
  SSK_CUDA(cudaEventRecord(g_chain.hop, producer));
  SSK_CUDA(cudaStreamWaitEvent(waiter, g_chain.hop, 0));
  return SSK_OK;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(); }

// interpolateCubic (OpenCV imgproc, A = -0.75) evaluated in float; compiled with -ffp-contract=off
static void cubic_coeffs_host(float x, float *c) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
  c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
  c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
  c[3] = 1.f - c[0] - c[1] - c[2];
}

static int sat_short(float v) {
  long r = lrintf(v);  // cvRound: round half to even
  return (int)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
}

// interpolateLanczos4 (OpenCV imgproc, imgwarp.cpp) evaluated as OpenCV does: sin / cos in double, coefficients in float
static void lanczos4_coeffs_host(float x, float *coeffs) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double kPi = 3.1415926535897932384626433832795;
  float sum = 0;
  const double y0 = -(x + 3) * kPi * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
  for (int i = 0; i < 8; i++) {
    const float y0_ = (x + 3 - i);
    if (std::fabs(y0_) >= 1e-6f) {
      const double y = -y0_ * kPi * 0.25;
      coeffs[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    } else {
      coeffs[i] = 1e30f;
    }
    sum += coeffs[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; i++) coeffs[i] *= sum;
}

struct TableStore {
  int device = -1;
  float4 *cubic = nullptr;
  short *itab = nullptr;
  float *lanczos = nullptr;
  short *lanczos_itab = nullptr;
};
static std::mutex g_tab_mutex;
static std::vector<TableStore> g_tabs;

int get_tables(Tables *t) {
  int dev = 0;
  SSK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  for (auto &s : g_tabs) {
    if (s.device == dev) { t->cubic = s.cubic; t->cubic_itab = s.itab; t->lanczos = s.lanczos; t->lanczos_itab = s.lanczos_itab; return SSK_OK; }
  }
  float coef[kInterTab][4];
  for (int i = 0; i < kInterTab; ++i) cubic_coeffs_host((float)i * (1.0f / kInterTab), coef[i]);
  // initInterTab2D(INTER_CUBIC, fixpt = true)
  std::vector<short> itab(kInterTab * kInterTab * 16);
  for (int i = 0; i < kInterTab; ++i) {
    for (int j = 0; j < kInterTab; ++j) {
      int it[4][4], isum = 0;
      for (int k1 = 0; k1 < 4; ++k1)
        for (int k2 = 0; k2 < 4; ++k2) {
          const float v = coef[i][k1] * coef[j][k2];
          isum += it[k1][k2] = sat_short(v * (float)kCoefScale);
        }
      if (isum != kCoefScale) {
        const int diff = isum - kCoefScale;
        int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
        for (int k1 = 2; k1 < 4; ++k1)
          for (int k2 = 2; k2 < 4; ++k2) {
            if (it[k1][k2] < it[mk1][mk2]) mk1 = k1, mk2 = k2;
            else if (it[k1][k2] > it[Mk1][Mk2]) Mk1 = k1, Mk2 = k2;
          }
        if (diff < 0) it[Mk1][Mk2] -= diff;
        else it[mk1][mk2] -= diff;
      }
      short *dst = &itab[(i * kInterTab + j) * 16];
      for (int k = 0; k < 16; ++k) dst[k] = (short)it[k / 4][k % 4];
    }
  }
  // initInterTab1D / initInterTab2D(INTER_LANCZOS4): 8 coefficients per 1/32 fraction, 8 x 8 fixed-point products with the sum
  // forced to 32768 on the largest (smallest) of the four central entries
  static float lz[kInterTab][8];
  for (int i = 0; i < kInterTab; ++i) lanczos4_coeffs_host((float)i * (1.0f / kInterTab), lz[i]);
  std::vector<short> lz_itab((size_t)kInterTab * kInterTab * 64);
  for (int i = 0; i < kInterTab; ++i) {
    for (int j = 0; j < kInterTab; ++j) {
      int it[8][8], isum = 0;
      for (int k1 = 0; k1 < 8; ++k1)
        for (int k2 = 0; k2 < 8; ++k2) {
          const float v = lz[i][k1] * lz[j][k2];
          isum += it[k1][k2] = sat_short(v * (float)kCoefScale);
        }
      if (isum != kCoefScale) {
        const int diff = isum - kCoefScale;
        int Mk1 = 4, Mk2 = 4, mk1 = 4, mk2 = 4;
        for (int k1 = 4; k1 < 6; ++k1)
          for (int k2 = 4; k2 < 6; ++k2) {
            if (it[k1][k2] < it[mk1][mk2]) mk1 = k1, mk2 = k2;
            else if (it[k1][k2] > it[Mk1][Mk2]) Mk1 = k1, Mk2 = k2;
          }
        if (diff < 0) it[Mk1][Mk2] -= diff;
        else it[mk1][mk2] -= diff;
      }
      short *dst = &lz_itab[(size_t)(i * kInterTab + j) * 64];
      for (int k = 0; k < 64; ++k) dst[k] = (short)it[k / 8][k % 8];
    }
  }
  TableStore s;
  s.device = dev;
  SSK_CUDA(cudaMalloc(&s.lanczos, sizeof(lz)));
  SSK_CUDA(cudaMalloc(&s.lanczos_itab, lz_itab.size() * sizeof(short)));
  SSK_CUDA(cudaMemcpy(s.lanczos, lz, sizeof(lz), cudaMemcpyHostToDevice));
  SSK_CUDA(cudaMemcpy(s.lanczos_itab, lz_itab.data(), lz_itab.size() * sizeof(short), cudaMemcpyHostToDevice));
  SSK_CUDA(cudaMalloc(&s.cubic, sizeof(coef)));
  SSK_CUDA(cudaMalloc(&s.itab, itab.size() * sizeof(short)));
  SSK_CUDA(cudaMemcpy(s.cubic, coef, sizeof(coef), cudaMemcpyHostToDevice));
  SSK_CUDA(cudaMemcpy(s.itab, itab.data(), itab.size() * sizeof(short), cudaMemcpyHostToDevice));
  g_tabs.push_back(s);
  t->cubic = s.cubic;
  t->cubic_itab = s.itab;
  t->lanczos = s.lanczos;
  t->lanczos_itab = s.lanczos_itab;
  return SSK_OK;
}

}  // namespace ssk

extern "C" {
const char *ssk_last_error(void) { return ssk::last_error().c_str(); }
int ssk_version(void) { return 100; }
int64_t ssk_kernel_launch_count(void) { return ssk::launch_count(); }
}
