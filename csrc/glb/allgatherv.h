// allgatherv (new-style): like allgather with a per-rank element count; rank i's
// block sits at the prefix-sum offset of the counts. Parity: gloo/allgatherv.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class AllgathervOptions : public detail::CollectiveOptionsBase {
 public:
  explicit AllgathervOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf) { elementSize = sizeof(T); in = std::move(buf); }
  template <typename T>
  void setInput(T* ptr, size_t n) { elementSize = sizeof(T); in = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf, std::vector<size_t> counts) {
    elementSize = sizeof(T);
    out = std::move(buf);
    elements = std::move(counts);
  }
  template <typename T>
  void setOutput(T* ptr, std::vector<size_t> counts) {
    size_t total = 0;
    for (auto c : counts) total += c;
    elementSize = sizeof(T);
    out = context->createUnboundBuffer(ptr, total * sizeof(T));
    elements = std::move(counts);
  }
  void setInputRaw(void* ptr, size_t n, size_t es) { elementSize = es; in = context->createUnboundBuffer(ptr, n * es); }
  void setOutputRaw(void* ptr, std::vector<size_t> counts, size_t es) {
    size_t total = 0;
    for (auto c : counts) total += c;
    elementSize = es;
    out = context->createUnboundBuffer(ptr, total * es);
    elements = std::move(counts);
  }

  std::unique_ptr<UnboundBuffer> in;
  std::unique_ptr<UnboundBuffer> out;
  std::vector<size_t> elements;  // per-rank element counts
  size_t elementSize = 0;
};

void allgatherv(AllgathervOptions& opts);

}  // namespace glb
