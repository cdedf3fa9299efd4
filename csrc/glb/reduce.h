// reduce (new-style): element-wise reduction of every rank's input delivered to
// `root` only. Ring reduce-scatter (shared with allreduce) followed by each rank
// sending its reduced chunk straight to the root. Parity: gloo/reduce.{h,cc}.
#pragma once

#include "glb/allreduce.h"

namespace glb {

class ReduceOptions : public detail::CollectiveOptionsBase {
 public:
  using Func = AllreduceOptions::Func;
  explicit ReduceOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf) { elements = buf->size / sizeof(T); elementSize = sizeof(T); in = std::move(buf); }
  template <typename T>
  void setInput(T* ptr, size_t n) { elements = n; elementSize = sizeof(T); in = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf) { elements = buf->size / sizeof(T); elementSize = sizeof(T); out = std::move(buf); }
  template <typename T>
  void setOutput(T* ptr, size_t n) { elements = n; elementSize = sizeof(T); out = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  void setInputRaw(void* ptr, size_t n, size_t es) { elements = n; elementSize = es; in = context->createUnboundBuffer(ptr, n * es); }
  void setOutputRaw(void* ptr, size_t n, size_t es) { elements = n; elementSize = es; out = context->createUnboundBuffer(ptr, n * es); }
  void setRoot(int r) { root = r; }
  void setReduceFunction(Func fn) { reduce = std::move(fn); }
  void setMaxSegmentSize(size_t s) { maxSegmentSize = s; }

  std::unique_ptr<UnboundBuffer> in;   // optional: in-place on out when absent
  std::unique_ptr<UnboundBuffer> out;  // required everywhere (scratch on non-root ranks)
  size_t elements = 0;
  size_t elementSize = 0;
  int root = -1;
  Func reduce;
  size_t maxSegmentSize = AllreduceOptions::kMaxSegmentSize;
};

void reduce(ReduceOptions& opts);

}  // namespace glb
