// reduce_scatter — both API generations.
//
//   reduce_scatter(ReduceScatterOptions&)   new-style function (the reference has
//       no new-style reduce_scatter; this is net-new): rank i receives
//       recvCounts[i] reduced elements. Direct algorithm: every rank sends each
//       peer the slice that peer owns and reduces what it receives — one step,
//       S(P-1)/P bytes per rank, the schedule that suits a full-bisection fabric.
//
//   ReduceScatterHalvingDoubling<T>         old-style class (reduce_scatter.cc):
//       recursive vector halving with bound buffers, binary blocks for
//       non-power-of-two P, user recvElems honoured by a final redistribution.
// Parity: gloo/reduce_scatter.h.
#pragma once

#include "glb/allreduce.h"

namespace glb {

class ReduceScatterOptions : public detail::CollectiveOptionsBase {
 public:
  using Func = AllreduceOptions::Func;
  explicit ReduceScatterOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

  template <typename T>
  void setInput(T* ptr, size_t n) { elementSize = sizeof(T); in = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  template <typename T>
  void setOutput(T* ptr, size_t n) { elementSize = sizeof(T); out = context->createUnboundBuffer(ptr, n * sizeof(T)); }
  void setInputRaw(void* ptr, size_t n, size_t es) { elementSize = es; in = context->createUnboundBuffer(ptr, n * es); }
  void setOutputRaw(void* ptr, size_t n, size_t es) { elementSize = es; out = context->createUnboundBuffer(ptr, n * es); }
  void setRecvCounts(std::vector<size_t> counts) { recvCounts = std::move(counts); }
  void setReduceFunction(Func fn) { reduce = std::move(fn); }

  std::unique_ptr<UnboundBuffer> in;   // sum(recvCounts) elements
  std::unique_ptr<UnboundBuffer> out;  // recvCounts[rank] elements
  std::vector<size_t> recvCounts;      // empty => equal split of the input
  size_t elementSize = 0;
  Func reduce;
};

void reduce_scatter(ReduceScatterOptions& opts);

}  // namespace glb
