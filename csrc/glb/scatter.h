// scatter (new-style): root holds P input buffers and sends in[i] to rank i.
// Parity: gloo/scatter.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class ScatterOptions : public detail::CollectiveOptionsBase {
 public:
  explicit ScatterOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}
  template <typename T>
  void setInputs(std::vector<std::unique_ptr<UnboundBuffer>> bufs) { elementSize = sizeof(T); in = std::move(bufs); }
  template <typename T>
  void setInputs(std::vector<T*> ptrs, size_t n) {
    elementSize = sizeof(T);
    in.clear();
    for (auto* p : ptrs) in.push_back(context->createUnboundBuffer(p, n * sizeof(T)));
  }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf) { elementSize = sizeof(T); out = std::move(buf); }
  template <typename T>
  void setOutput(T* ptr, size_t n) { elementSize = sizeof(T); out = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  void setInputsRaw(const std::vector<void*>& ptrs, size_t bytes) {
    in.clear();
    for (auto* p : ptrs) in.push_back(context->createUnboundBuffer(p, bytes));
    if (!elementSize) elementSize = 1;
  }
  void setOutputRaw(void* ptr, size_t bytes) { out = context->createUnboundBuffer(ptr, bytes); if (!elementSize) elementSize = 1; }
  void setRoot(int r) { root = r; }

  std::vector<std::unique_ptr<UnboundBuffer>> in;  // root only, one per rank
  std::unique_ptr<UnboundBuffer> out;
  size_t elementSize = 0;
  int root = -1;
};

void scatter(ScatterOptions& opts);

}  // namespace glb
