// Multi-GPU epilogue of the stacking path in the C ABI (ssk_acc_reduce(h, ncclComm_t, root)).
//
// The reference is single-process; what its pipeline holds at the end of a run is one c_frame_accumulation
// (c_frame_accumulation.h:14-63).  With frames sharded over ranks every rank ends with a local accumulator, and
//   c_weigthed_average:  A = sum_g(W_g A_g) / sum_g(W_g),  W = sum_g W_g      (running mean <-> sum form, c_frame_accumulation.cc:20-129)
//   c_bayer_average:     acc = sum_g acc_g,  cntr = sum_g cntr_g             (already sums, c_frame_accumulation.cc:988-1126)
// so the combine is: sum form in place, ONE ncclReduce group (values, weights, frame count) to the root over NVLink,
// back to the running mean.  No data-path collective exists anywhere else on the path.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, or the copy the process already loaded): libssk.so carries no
// link-time dependency on it and single-GPU users never touch it.  The communicator is the caller's (a C++ host creates
// it with ncclCommInitRank); ssk_nccl_* are thin conveniences for hosts that do not link NCCL themselves (ctypes).
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <nccl.h>
#include <mutex>
#include "ssk_engine.cuh"

using namespace ssk;

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *lib = nullptr;
    if (const char *p = getenv("SSK_NCCL_LIB")) lib = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    // a copy already mapped into the process (e.g. the one torch bundles) wins over the system library
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char *n) { return dlsym(lib, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Reduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
    if (!api.ok) api.why = "libnccl.so.2 lacks an expected symbol";
  });
  return api;
}

int nccl_fail(ncclResult_t r, const char *what) {
  set_error(std::string("NCCL: ") + what + ": " + nccl().GetErrorString(r));
  return SSK_ERR_NCCL;
}

#define SSK_NCCL(expr)                                     \
  do {                                                     \
    ncclResult_t _r = (expr);                              \
    if (_r != ncclSuccess) return nccl_fail(_r, #expr);    \
  } while (0)

int require_nccl() {
  if (nccl().ok) return SSK_OK;
  set_error("NCCL is not available: " + nccl().why);
  return SSK_ERR_NCCL;
}

// frames_total: device int that holds this rank's accumulated-frame count on entry and the total on the root on return
int reduce_accumulator(Acc &a, void *comm_, int root, int *d_frames, cudaStream_t s) {
  ncclComm_t comm = static_cast<ncclComm_t>(comm_);
  const int64_t npix = (int64_t)a.rows * a.cols;
  const bool bayer = a.kind == SSK_ACC_BAYER_AVERAGE;
  const size_t nacc = (size_t)npix * (bayer ? 3 : a.cn), nw = (size_t)npix * (bayer ? 3 : 1);
  if (!bayer) { if (int e = launch_acc_sum_form(a.acc.as<float>(), a.wacc.as<float>(), npix, a.cn, 1, s)) return e; }
  NcclApi &n = nccl();
  SSK_NCCL(n.GroupStart());
  SSK_NCCL(n.Reduce(a.acc.p, a.acc.p, nacc, ncclFloat32, ncclSum, root, comm, s));
  SSK_NCCL(n.Reduce(a.wacc.p, a.wacc.p, nw, ncclFloat32, ncclSum, root, comm, s));
  SSK_NCCL(n.Reduce(d_frames, d_frames, 1, ncclInt32, ncclSum, root, comm, s));
  SSK_NCCL(n.GroupEnd());
  // every rank returns to the running mean: the root of the total, the others of what they still hold
  if (!bayer) { if (int e = launch_acc_sum_form(a.acc.as<float>(), a.wacc.as<float>(), npix, a.cn, 0, s)) return e; }
  return SSK_OK;
}

}  // namespace

// defined in ssk_stack.cu: the pipeline's stream, accumulator and device-side frame counter
int stack_reduce_parts(ssk_stack *h, cudaStream_t *stream, ssk_acc **acc, int **d_counter);

extern "C" {

int ssk_nccl_get_unique_id(void *id128) {
  SSK_REQUIRE(id128, "null argument");
  if (int e = require_nccl()) return e;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  SSK_NCCL(nccl().GetUniqueId(static_cast<ncclUniqueId *>(id128)));
  return SSK_OK;
}

int ssk_nccl_comm_create(const void *id128, int nranks, int rank, void **comm) {
  SSK_REQUIRE(id128 && comm && nranks >= 1 && rank >= 0 && rank < nranks, "ssk_nccl_comm_create: bad argument");
  if (int e = require_nccl()) return e;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  SSK_NCCL(nccl().CommInitRank(&c, nranks, id, rank));
  *comm = c;
  return SSK_OK;
}

int ssk_nccl_comm_destroy(void *comm) {
  if (!comm) return SSK_OK;
  if (int e = require_nccl()) return e;
  SSK_NCCL(nccl().CommDestroy(static_cast<ncclComm_t>(comm)));
  return SSK_OK;
}

int ssk_acc_reduce(ssk_acc *h, void *nccl_comm, int root) {
  SSK_REQUIRE(h && nccl_comm && h->a.acc.p, "ssk_acc_reduce: null argument / empty accumulator");
  if (int e = require_nccl()) return e;
  Acc &a = h->a;
  static thread_local DevBuf d_n;
  if (int e = d_n.ensure(sizeof(int))) return e;
  SSK_CUDA(cudaMemcpyAsync(d_n.p, &a.frames, sizeof(int), cudaMemcpyHostToDevice, a.stream));
  if (int e = reduce_accumulator(a, nccl_comm, root, d_n.as<int>(), a.stream)) return e;
  int total = a.frames;
  SSK_CUDA(cudaMemcpyAsync(&total, d_n.p, sizeof(int), cudaMemcpyDeviceToHost, a.stream));
  SSK_CUDA(cudaStreamSynchronize(a.stream));
  a.frames = total;     // the root holds the sum over the ranks, every other rank its own count (ncclReduce leaves it)
  return SSK_OK;
}

int ssk_stack_reduce(ssk_stack *h, void *nccl_comm, int root) {
  SSK_REQUIRE(h && nccl_comm, "ssk_stack_reduce: null argument");
  if (int e = require_nccl()) return e;
  if (int e = ssk_stack_flush(h)) return e;
  cudaStream_t s = nullptr;
  ssk_acc *acc = nullptr;
  int *d_counter = nullptr;
  if (int e = stack_reduce_parts(h, &s, &acc, &d_counter)) return e;
  SSK_REQUIRE(acc->a.acc.p, "ssk_stack_reduce: set_reference must be called first");
  // the device-side frame counter is reduced in place, so that accumulated_frames() / compute() on the root see the total
  if (int e = reduce_accumulator(acc->a, nccl_comm, root, d_counter, s)) return e;
  return ssk_stack_sync(h);
}

}  // extern "C"
