// allreduce (new-style API): every rank contributes one or more input buffers;
// afterwards every output buffer on every rank holds the element-wise reduction.
//
// Algorithms
//   RING    bandwidth-optimal ring: reduce-scatter then allgather, each chunk
//           streamed in segments of at most maxSegmentSize bytes with two receives
//           in flight so the CPU reduction overlaps the wire.
//   BCUBE   mixed-radix recursive halving/doubling. P is factorised (2·2·… for
//           powers of two => classic halving-doubling; any other P uses its prime
//           factors, a prime P degenerates to a direct exchange) so non-power-of-two
//           sizes need no special casing. 2·Σ(f_i − 1) messages per rank instead of
//           the ring's 2(P − 1) steps.
// Parity: gloo/allreduce.{h,cc}.
#pragma once

#include "glb/collectives_common.h"
#include "glb/math.h"

namespace glb {

class AllreduceOptions : public detail::CollectiveOptionsBase {
 public:
  using Func = std::function<void(void*, const void*, const void*, size_t)>;

  enum Algorithm { UNSPECIFIED = 0, RING = 1, BCUBE = 2 };

  explicit AllreduceOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

  void setAlgorithm(Algorithm a) { algorithm = a; }

  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf) {
    std::vector<std::unique_ptr<UnboundBuffer>> v;
    v.push_back(std::move(buf));
    setInputs<T>(std::move(v));
  }
  template <typename T>
  void setInputs(std::vector<std::unique_ptr<UnboundBuffer>> bufs) {
    GLB_ENFORCE(!bufs.empty());
    elements = bufs[0]->size / sizeof(T);
    elementSize = sizeof(T);
    in = std::move(bufs);
  }
  template <typename T>
  void setInput(T* ptr, size_t n) { setInputs<T>(&ptr, 1, n); }
  template <typename T>
  void setInputs(std::vector<T*> ptrs, size_t n) { setInputs<T>(ptrs.data(), ptrs.size(), n); }
  template <typename T>
  void setInputs(T** ptrs, size_t len, size_t n) {
    elements = n;
    elementSize = sizeof(T);
    in.clear();
    for (size_t i = 0; i < len; i++) in.push_back(context->createUnboundBuffer(ptrs[i], n * sizeof(T)));
  }

  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf) {
    std::vector<std::unique_ptr<UnboundBuffer>> v;
    v.push_back(std::move(buf));
    setOutputs<T>(std::move(v));
  }
  template <typename T>
  void setOutputs(std::vector<std::unique_ptr<UnboundBuffer>> bufs) {
    GLB_ENFORCE(!bufs.empty());
    elements = bufs[0]->size / sizeof(T);
    elementSize = sizeof(T);
    out = std::move(bufs);
  }
  template <typename T>
  void setOutput(T* ptr, size_t n) { setOutputs<T>(&ptr, 1, n); }
  template <typename T>
  void setOutputs(std::vector<T*> ptrs, size_t n) { setOutputs<T>(ptrs.data(), ptrs.size(), n); }
  template <typename T>
  void setOutputs(T** ptrs, size_t len, size_t n) {
    elements = n;
    elementSize = sizeof(T);
    out.clear();
    for (size_t i = 0; i < len; i++) out.push_back(context->createUnboundBuffer(ptrs[i], n * sizeof(T)));
  }

  // Type-erased variants (used by the language bindings).
  void setInputsRaw(const std::vector<void*>& ptrs, size_t n, size_t elemSize);
  void setOutputsRaw(const std::vector<void*>& ptrs, size_t n, size_t elemSize);

  void setReduceFunction(Func fn) { reduce = std::move(fn); }
  void setMaxSegmentSize(size_t s) { maxSegmentSize = s; }

  static constexpr size_t kMaxSegmentSize = 1024 * 1024;

  Algorithm algorithm = UNSPECIFIED;
  std::vector<std::unique_ptr<UnboundBuffer>> in;
  std::vector<std::unique_ptr<UnboundBuffer>> out;
  size_t elements = 0;
  size_t elementSize = 0;
  Func reduce;
  size_t maxSegmentSize = kMaxSegmentSize;
};

void allreduce(const AllreduceOptions& opts);

namespace detail {
// Ring reduce-scatter over `buf` (elements * elementSize bytes), shared with reduce()
// and reduce_scatter(): on return rank r holds the fully reduced chunk
// subRange({0, elements}, P, (r + 1) % P). Uses slots [slot, slot + 1).
void ringReduceScatter(const std::shared_ptr<Context>& context, UnboundBuffer* buf, size_t elements,
                       size_t elementSize, const AllreduceOptions::Func& reduce, size_t maxSegmentSize,
                       uint64_t slot, std::chrono::milliseconds timeout);
// Mixed-radix factorisation used by BCUBE (exposed for tests).
std::vector<int> factorize(int n);
}  // namespace detail

}  // namespace glb
