// alltoall (new-style): input and output are P equal chunks; chunk j of rank i's
// input lands in chunk i of rank j's output. All P-1 sends and receives are posted
// at once (rank+i / rank-i pairing spreads the load). Parity: gloo/alltoall.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class AlltoallOptions : public detail::CollectiveOptionsBase {
 public:
  explicit AlltoallOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

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

  std::unique_ptr<UnboundBuffer> in;
  std::unique_ptr<UnboundBuffer> out;
  size_t elementSize = 0;
};

void alltoall(AlltoallOptions& opts);

}  // namespace glb
