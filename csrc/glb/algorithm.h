// Base classes of the old-style (class per algorithm) API: Algorithm,
// ReductionFunction<T>, LocalOp<T>. Parity: gloo/algorithm.{h,cc}.
#pragma once

#include <memory>

#include "glb/common/logging.h"
#include "glb/context.h"
#include "glb/math.h"
#include "glb/types.h"

namespace glb {

// Messages up to this many bytes are latency-bound: the CUDA algorithms pick the
// one-shot kernel below it (the reference uses the same constant to choose host
// vs device reduction, algorithm.cc:16).
constexpr size_t kOnDeviceThreshold = 256 * 1024;

class Algorithm {
 public:
  explicit Algorithm(const std::shared_ptr<Context>& context);
  virtual ~Algorithm() noexcept(false);
  virtual void run() = 0;

 protected:
  std::shared_ptr<Context> context_;
  const int contextRank_;
  const int contextSize_;

  std::unique_ptr<transport::Pair>& getPair(int i);
  // Ring neighbours.
  std::unique_ptr<transport::Pair>& getLeftPair();
  std::unique_ptr<transport::Pair>& getRightPair();
};

// Kept as an alias: call sites in the reference spell it ReductionType.
using ReductionType = ReduceOp;

template <typename T>
class ReductionFunction {
 public:
  using Function = void(T*, const T*, size_t n);

  static const ReductionFunction<T>* sum;
  static const ReductionFunction<T>* product;
  static const ReductionFunction<T>* min;
  static const ReductionFunction<T>* max;
  static const ReductionFunction<T>* get(ReduceOp op);

  ReductionFunction(ReduceOp type, ReduceFn fn) : type_(type), fn_(fn) {}
  ReduceOp type() const { return type_; }
  // x[i] = x[i] (op) y[i]
  void call(T* x, const T* y, size_t n) const { fn_(x, x, y, n); }
  ReduceFn raw() const { return fn_; }

 protected:
  ReduceOp type_;
  ReduceFn fn_;
};

template <typename T>
const ReductionFunction<T>* ReductionFunction<T>::sum = new ReductionFunction<T>(ReduceOp::SUM, &::glb::sum<T>);
template <typename T>
const ReductionFunction<T>* ReductionFunction<T>::product =
    new ReductionFunction<T>(ReduceOp::PRODUCT, &::glb::product<T>);
template <typename T>
const ReductionFunction<T>* ReductionFunction<T>::min = new ReductionFunction<T>(ReduceOp::MIN, &::glb::min<T>);
template <typename T>
const ReductionFunction<T>* ReductionFunction<T>::max = new ReductionFunction<T>(ReduceOp::MAX, &::glb::max<T>);

template <typename T>
const ReductionFunction<T>* ReductionFunction<T>::get(ReduceOp op) {
  switch (op) {
    case ReduceOp::SUM: return sum;
    case ReduceOp::PRODUCT: return product;
    case ReduceOp::MIN: return min;
    case ReduceOp::MAX: return max;
    default: return nullptr;
  }
}

// Local operation (intra-process, possibly spanning several GPUs): runAsync
// enqueues, wait blocks until done.
template <typename T>
class LocalOp {
 public:
  virtual ~LocalOp() noexcept(false) {}
  virtual void runAsync() = 0;
  virtual void wait() = 0;
  inline void run() {
    runAsync();
    wait();
  }
};

}  // namespace glb
