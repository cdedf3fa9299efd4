// allgather (new-style): rank i's input lands at output[i*n, (i+1)*n) everywhere.
// Ring schedule, each block split in two halves so two sends and two receives are
// always in flight. In-place when no input is given (own block already in output).
// Parity: gloo/allgather.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class AllgatherOptions : public detail::CollectiveOptionsBase {
 public:
  explicit AllgatherOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

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

void allgather(AllgatherOptions& opts);

}  // namespace glb
