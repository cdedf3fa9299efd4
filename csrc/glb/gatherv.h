// gatherv (new-style): gather with per-rank element counts. Parity: gloo/gatherv.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class GathervOptions : public detail::CollectiveOptionsBase {
 public:
  explicit GathervOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}
  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf) { elementSize = sizeof(T); in = std::move(buf); }
  template <typename T>
  void setInput(T* ptr, size_t n) { elementSize = sizeof(T); in = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf, std::vector<size_t> counts) {
    elementSize = sizeof(T);
    out = std::move(buf);
    elementsPerRank = std::move(counts);
  }
  template <typename T>
  void setOutput(T* ptr, std::vector<size_t> counts) { setOutputRaw(ptr, std::move(counts), sizeof(T)); }
  void setInputRaw(void* ptr, size_t n, size_t es) { elementSize = es; in = context->createUnboundBuffer(ptr, n * es); }
  void setOutputRaw(void* ptr, std::vector<size_t> counts, size_t es) {
    size_t total = 0;
    for (auto c : counts) total += c;
    elementSize = es;
    out = context->createUnboundBuffer(ptr, total * es);
    elementsPerRank = std::move(counts);
  }
  void setRoot(int r) { root = r; }

  std::unique_ptr<UnboundBuffer> in;
  std::unique_ptr<UnboundBuffer> out;       // root only
  std::vector<size_t> elementsPerRank;      // root only
  size_t elementSize = 0;
  int root = -1;
};

void gatherv(GathervOptions& opts);

}  // namespace glb
