// Chrome-trace ("chrome://tracing" / Perfetto JSON) event sink for collective calls.
// GLB_TRACE_FILE=<path> turns it on: every traced scope becomes one complete event
// ({"ph":"X"}) with wall-clock start and duration, process id and thread id; "%r" in the path
// is replaced by the process id so that one-process-per-rank jobs write one file per rank.
// Events are buffered and written when the process exits (or by flush()). Off by default and
// then a traced scope costs one relaxed load. The reference has no tracing hooks (SURVEY
// section 5); the CUDA collectives additionally emit NVTX ranges (cuda/trace.h).
#pragma once

#include <cstddef>
#include <cstdint>

namespace glb {
namespace trace {

bool enabled();
uint64_t nowNs();
void record(const char* name, uint64_t startNs, uint64_t endNs);
// Write what has been recorded so far (also runs at exit). Returns the number of events written.
size_t flush();

class Scope {
 public:
  explicit Scope(const char* name) : name_(name), start_(enabled() ? nowNs() : 0) {}
  ~Scope() {
    if (start_ != 0) record(name_, start_, nowNs());
  }
  Scope(const Scope&) = delete;
  Scope& operator=(const Scope&) = delete;

 private:
  const char* name_;
  uint64_t start_;
};

}  // namespace trace
}  // namespace glb

#define GLB_HOST_TRACE_CONCAT2(a, b) a##b
#define GLB_HOST_TRACE_CONCAT(a, b) GLB_HOST_TRACE_CONCAT2(a, b)
#define GLB_HOST_TRACE(name) ::glb::trace::Scope GLB_HOST_TRACE_CONCAT(glb_host_trace_, __LINE__)(name)
