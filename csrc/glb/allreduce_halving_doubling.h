// AllreduceHalvingDoubling<T> (old-style): recursive vector halving with distance
// doubling (reduce-scatter, lg P steps) followed by the mirrored allgather; 2·lg P
// steps, 2·S bytes. Non-power-of-two P: binary blocks (binary_blocks.h; the reference's
// scheme, allreduce_halving_doubling.h:39-64): blocks of decreasing power-of-two size run
// the halving phase concurrently and are chained for the cross-block reduction, so no rank
// idles. GLB_HD_FOLD=1 selects the simpler alternative instead: the P - 2^k surplus ranks
// fold their vector onto a partner first and get the result back afterwards (one extra
// full-vector hop on each side).
// Parity: gloo/allreduce_halving_doubling.h:38-415.
#pragma once

#include "glb/binary_blocks.h"
#include "glb/mixed_radix.h"

namespace glb {

template <typename T>
class AllreduceHalvingDoubling : public Algorithm {
 public:
  AllreduceHalvingDoubling(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs,
                           const size_t count, const ReductionFunction<T>* fn = ReductionFunction<T>::sum)
      : Algorithm(context), ptrs_(ptrs), count_(count), bytes_(count * sizeof(T)), fn_(fn) {
    GLB_ENFORCE(!ptrs_.empty());
    if (contextSize_ == 1) return;
    const int core = detail::largestPow2AtMost(contextSize_);
    if (core != contextSize_ && !envFlag("HD_FOLD", false)) {
      blocks_.reset(new detail::BinaryBlocksAllreduce<T>(context_, ptrs_[0], count_, fn_));
      return;
    }
    std::vector<int> factors(log2ceil(static_cast<uint32_t>(core)), 2);
    engine_.reset(new detail::MixedRadix<T>(this, context_, ptrs_[0], count_, fn_, factors, core, true));
  }

  void run() override {
    if (count_ == 0) return;
    for (size_t i = 1; i < ptrs_.size(); i++) fn_->call(ptrs_[0], ptrs_[i], count_);
    if (blocks_) {
      blocks_->run();
    } else if (engine_) {
      engine_->foldIn();
      engine_->reduceScatter();
      engine_->allgather();
      engine_->foldOut(0, count_);
    }
    for (size_t i = 1; i < ptrs_.size(); i++) std::memcpy(ptrs_[i], ptrs_[0], bytes_);
  }

 protected:
  std::vector<T*> ptrs_;
  const size_t count_;
  const size_t bytes_;
  const ReductionFunction<T>* fn_;
  std::unique_ptr<detail::MixedRadix<T>> engine_;
  std::unique_ptr<detail::BinaryBlocksAllreduce<T>> blocks_;
};

}  // namespace glb
