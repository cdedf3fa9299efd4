// gather (new-style): root collects every rank's input at output[i*n]. Non-root
// ranks send directly; the root posts all receives up front. Parity: gloo/gather.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class GatherOptions : public detail::CollectiveOptionsBase {
 public:
  explicit GatherOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}
  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf) { elementSize = sizeof(T); in = std::move(buf); }
  template <typename T>
  void setInput(T* ptr, size_t n) { elementSize = sizeof(T); in = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf) { elementSize = sizeof(T); out = std::move(buf); }
  template <typename T>
  void setOutput(T* ptr, size_t n) { elementSize = sizeof(T); out = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  void setInputRaw(void* ptr, size_t bytes) { in = context->createUnboundBuffer(ptr, bytes); if (!elementSize) elementSize = 1; }
  void setOutputRaw(void* ptr, size_t bytes) { out = context->createUnboundBuffer(ptr, bytes); if (!elementSize) elementSize = 1; }
  void setRoot(int r) { root = r; }

  std::unique_ptr<UnboundBuffer> in;
  std::unique_ptr<UnboundBuffer> out;  // root only
  size_t elementSize = 0;
  int root = -1;
};

void gather(GatherOptions& opts);

}  // namespace glb
