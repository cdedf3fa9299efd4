#include "glb/algorithm.h"

#include "glb/common/logging.h"

namespace glb {

Algorithm::Algorithm(const std::shared_ptr<Context>& context)
    : context_(context), contextRank_(context->rank), contextSize_(context->size) {}

Algorithm::~Algorithm() noexcept(false) {}

std::unique_ptr<transport::Pair>& Algorithm::getPair(int i) { return context_->getPair(i); }

std::unique_ptr<transport::Pair>& Algorithm::getLeftPair() {
  int rank = (contextSize_ + contextRank_ - 1) % contextSize_;
  GLB_ENFORCE(context_->getPair(rank), "pair missing (index ", rank, ")");
  return context_->getPair(rank);
}

std::unique_ptr<transport::Pair>& Algorithm::getRightPair() {
  int rank = (contextRank_ + 1) % contextSize_;
  GLB_ENFORCE(context_->getPair(rank), "pair missing (index ", rank, ")");
  return context_->getPair(rank);
}

}  // namespace glb
