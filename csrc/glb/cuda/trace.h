// NVTX ranges around every CUDA collective (visible in Nsight Systems / Compute
// timelines). NVTX v3 is header-only and binds to the tool's injection library at run
// time, so this costs nothing when no profiler is attached. The reference has no tracing
// hooks at all; GLB_NVTX=0 disables them.
#pragma once

#include <nvtx3/nvToolsExt.h>

#include "glb/common/trace.h"
#include "glb/common/utils.h"

namespace glb {
namespace cuda {

inline bool nvtxEnabled() {
  static const bool on = envFlag("NVTX", true);
  return on;
}

class TraceRange {
 public:
  explicit TraceRange(const char* name) : active_(nvtxEnabled()), host_(name) {
    if (active_) nvtxRangePushA(name);
  }
  ~TraceRange() {
    if (active_) nvtxRangePop();
  }
  TraceRange(const TraceRange&) = delete;
  TraceRange& operator=(const TraceRange&) = delete;

 private:
  bool active_;
  ::glb::trace::Scope host_;  // the same scope as a chrome-trace event when GLB_TRACE_FILE is set (host-side enqueue time)
};

}  // namespace cuda
}  // namespace glb

#define GLB_TRACE_CONCAT2(a, b) a##b
#define GLB_TRACE_CONCAT(a, b) GLB_TRACE_CONCAT2(a, b)
#define GLB_TRACE_RANGE(name) ::glb::cuda::TraceRange GLB_TRACE_CONCAT(glb_trace_, __LINE__)(name)
