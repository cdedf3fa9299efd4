// AllreduceBcube<T> (old-style): hypercube allreduce with base B = context->base.
// Each of the ceil(log_B P) steps exchanges with up to B-1 peers at once; P need
// not be a power of B — the radix of each step is the largest divisor of what is
// left that does not exceed B (a prime remainder larger than B becomes one direct
// step). 2·log_B P steps. Parity: gloo/allreduce_bcube.h:265-702.
#pragma once

#include "glb/mixed_radix.h"

namespace glb {

template <typename T>
class AllreduceBcube : public Algorithm {
 public:
  AllreduceBcube(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs, const size_t count,
                 const ReductionFunction<T>* fn = ReductionFunction<T>::sum)
      : Algorithm(context), ptrs_(ptrs), count_(count), bytes_(count * sizeof(T)), fn_(fn) {
    GLB_ENFORCE(!ptrs_.empty());
    GLB_ENFORCE_GE(context->base, 2, "bcube base must be at least 2");
    if (contextSize_ == 1) return;
    auto factors = detail::radixFactors(contextSize_, context->base);
    engine_.reset(new detail::MixedRadix<T>(this, context_, ptrs_[0], count_, fn_, factors, contextSize_, true));
  }

  void run() override {
    if (count_ == 0) return;
    for (size_t i = 1; i < ptrs_.size(); i++) fn_->call(ptrs_[0], ptrs_[i], count_);
    if (engine_) {
      engine_->reduceScatter();
      engine_->allgather();
    }
    for (size_t i = 1; i < ptrs_.size(); i++) std::memcpy(ptrs_[i], ptrs_[0], bytes_);
  }

  int steps() const { return engine_ ? engine_->numSteps() : 0; }

 protected:
  std::vector<T*> ptrs_;
  const size_t count_;
  const size_t bytes_;
  const ReductionFunction<T>* fn_;
  std::unique_ptr<detail::MixedRadix<T>> engine_;
};

}  // namespace glb
