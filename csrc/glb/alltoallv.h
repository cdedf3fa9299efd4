// alltoallv (new-style): alltoall with per-destination element counts; offsets are
// the prefix sums of the counts. Parity: gloo/alltoallv.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class AlltoallvOptions : public detail::CollectiveOptionsBase {
 public:
  explicit AlltoallvOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

  template <typename T>
  void setInput(T* ptr, std::vector<int64_t> counts) { setInputRaw(ptr, std::move(counts), sizeof(T)); }
  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf, std::vector<int64_t> counts) {
    setCounts(std::move(counts), sizeof(T), inOffsetPerRank, inLengthPerRank);
    in = std::move(buf);
  }
  template <typename T>
  void setOutput(T* ptr, std::vector<int64_t> counts) { setOutputRaw(ptr, std::move(counts), sizeof(T)); }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf, std::vector<int64_t> counts) {
    setCounts(std::move(counts), sizeof(T), outOffsetPerRank, outLengthPerRank);
    out = std::move(buf);
  }
  void setInputRaw(void* ptr, std::vector<int64_t> counts, size_t es) {
    size_t total = setCounts(std::move(counts), es, inOffsetPerRank, inLengthPerRank);
    in = context->createUnboundBuffer(ptr, total);
  }
  void setOutputRaw(void* ptr, std::vector<int64_t> counts, size_t es) {
    size_t total = setCounts(std::move(counts), es, outOffsetPerRank, outLengthPerRank);
    out = context->createUnboundBuffer(ptr, total);
  }

  std::unique_ptr<UnboundBuffer> in;
  std::unique_ptr<UnboundBuffer> out;
  std::vector<size_t> inOffsetPerRank, inLengthPerRank;    // bytes
  std::vector<size_t> outOffsetPerRank, outLengthPerRank;  // bytes
  size_t elementSize = 0;

 private:
  size_t setCounts(std::vector<int64_t> counts, size_t es, std::vector<size_t>& off, std::vector<size_t>& len) {
    GLB_ENFORCE_EQ(static_cast<int>(counts.size()), context->size, "alltoallv: need one count per rank");
    GLB_ENFORCE(elementSize == 0 || elementSize == es, "alltoallv: element size mismatch");
    elementSize = es;
    off.clear();
    len.clear();
    size_t o = 0;
    for (auto c : counts) {
      GLB_ENFORCE_GE(c, 0, "alltoallv: negative count");
      off.push_back(o);
      len.push_back(static_cast<size_t>(c) * es);
      o += static_cast<size_t>(c) * es;
    }
    return o;
  }
};

void alltoallv(AlltoallvOptions& opts);

}  // namespace glb
