// NVTX ranges of the per-frame loop's stages (the reference logs per-stage times through its
// c_image_processing_pipeline timers; here the stages are enqueued asynchronously, so besides the CUDA-event times of
// ssk_stack_stage_times the host-side enqueue of every stage is bracketed by an NVTX range that a timeline tool
// correlates with the kernels it launched).  Header-only NVTX v3: a no-op (one branch) unless a tool injected itself.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace ssk {

// One open range at a time: next() closes the previous stage and opens the named one; the destructor closes the last.
struct NvtxStage {
  bool open = false;
  void next(const char *name) {
    if (open) nvtxRangePop();
    nvtxRangePushA(name);
    open = true;
  }
  void end() {
    if (open) nvtxRangePop();
    open = false;
  }
  ~NvtxStage() { end(); }
};

struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

}  // namespace ssk
