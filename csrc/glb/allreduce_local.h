// AllreduceLocal<T>: reduce + broadcast across the local pointers only.
// Parity: gloo/allreduce_local.{h,cc}.
#pragma once

#include <cstring>
#include <vector>

#include "glb/algorithm.h"

namespace glb {

template <typename T>
class AllreduceLocal : public Algorithm {
 public:
  AllreduceLocal(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs, const size_t count,
                 const ReductionFunction<T>* fn = ReductionFunction<T>::sum)
      : Algorithm(context), ptrs_(ptrs), count_(count), bytes_(count * sizeof(T)), fn_(fn) {}

  void run() override {
    for (size_t i = 1; i < ptrs_.size(); i++) fn_->call(ptrs_[0], ptrs_[i], count_);
    for (size_t i = 1; i < ptrs_.size(); i++) std::memcpy(ptrs_[i], ptrs_[0], bytes_);
  }

 protected:
  std::vector<T*> ptrs_;
  const size_t count_;
  const size_t bytes_;
  const ReductionFunction<T>* fn_;
};

}  // namespace glb
