// ReduceScatterHalvingDoubling<T> (old-style): recursive-halving reduce-scatter
// followed by a distribution phase that honours the caller's recvElems (how many
// reduced elements each rank wants): after the halving every rank owns an even
// 1/P block; blocks are then cut along the recvElems boundaries and shipped to
// their final owners. The result lands at the start of ptrs[0].
// Non-power-of-two P uses the prime factorisation of P as radices.
// Parity: gloo/reduce_scatter.h:22-505.
#pragma once

#include <numeric>

#include "glb/mixed_radix.h"

namespace glb {

template <typename T>
class ReduceScatterHalvingDoubling : public Algorithm {
 public:
  ReduceScatterHalvingDoubling(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs,
                               const size_t count, const std::vector<int> recvElems,
                               const ReductionFunction<T>* fn = ReductionFunction<T>::sum)
      : Algorithm(context), ptrs_(ptrs), count_(count), fn_(fn), recvElems_(recvElems) {
    GLB_ENFORCE(!ptrs_.empty());
    GLB_ENFORCE_EQ(static_cast<int>(recvElems_.size()), contextSize_, "need one recv count per rank");
    size_t total = 0;
    for (int e : recvElems_) {
      GLB_ENFORCE_GE(e, 0);
      total += static_cast<size_t>(e);
    }
    GLB_ENFORCE_EQ(total, count_, "recvElems must add up to count");
    if (contextSize_ == 1) return;
    auto factors = detail::radixFactors(contextSize_, 2);  // 2s first, then the odd primes
    engine_.reset(new detail::MixedRadix<T>(this, context_, ptrs_[0], count_, fn_, factors, contextSize_, false));

    // Distribution plan: intersect what each rank owns with what each rank wants.
    std::vector<size_t> want(contextSize_ + 1, 0);
    for (int i = 0; i < contextSize_; i++) want[i + 1] = want[i] + static_cast<size_t>(recvElems_[i]);
    result_.resize(std::max<size_t>(1, static_cast<size_t>(recvElems_[contextRank_])));
    const int slot = context_->nextSlot();
    const detail::Range mine = engine_->ownedBlock();
    for (int j = 0; j < contextSize_; j++) {
      // what I own that j wants
      size_t lo = std::max(mine.off, want[j]), hi = std::min(mine.off + mine.len, want[j + 1]);
      if (lo < hi) {
        if (j == contextRank_) {
          localCopy_ = Piece{lo, hi - lo, lo - want[j]};
        } else {
          Out o;
          o.piece = Piece{lo, hi - lo, lo - want[j]};
          o.buf = getPair(j)->createSendBuffer(slot, ptrs_[0], count_ * sizeof(T));
          outs_.push_back(std::move(o));
        }
      }
      // what j owns that I want
      if (j != contextRank_) {
        const detail::Range theirs = engine_->ownedBlockOf(j);
        size_t l2 = std::max(theirs.off, want[contextRank_]);
        size_t h2 = std::min(theirs.off + theirs.len, want[contextRank_ + 1]);
        if (l2 < h2) {
          ins_.push_back(getPair(j)->createRecvBuffer(slot, result_.data(), result_.size() * sizeof(T)));
        }
      }
    }
  }

  void run() override {
    if (count_ == 0) return;
    for (size_t i = 1; i < ptrs_.size(); i++) fn_->call(ptrs_[0], ptrs_[i], count_);
    if (!engine_) return;  // single rank: the whole vector is already in place
    engine_->reduceScatter();
    for (auto& o : outs_) o.buf->send(o.piece.srcOff * sizeof(T), o.piece.len * sizeof(T), o.piece.dstOff * sizeof(T));
    if (localCopy_.len > 0) {
      std::memcpy(result_.data() + localCopy_.dstOff, ptrs_[0] + localCopy_.srcOff, localCopy_.len * sizeof(T));
    }
    for (auto& b : ins_) b->waitRecv();
    for (auto& o : outs_) o.buf->waitSend();
    if (recvElems_[contextRank_] > 0) {
      std::memcpy(ptrs_[0], result_.data(), static_cast<size_t>(recvElems_[contextRank_]) * sizeof(T));
    }
  }

 protected:
  struct Piece {
    size_t srcOff = 0, len = 0, dstOff = 0;
  };
  struct Out {
    Piece piece;
    std::unique_ptr<transport::Buffer> buf;
  };
  std::vector<T*> ptrs_;
  const size_t count_;
  const ReductionFunction<T>* fn_;
  std::vector<int> recvElems_;
  std::unique_ptr<detail::MixedRadix<T>> engine_;
  std::vector<T> result_;
  Piece localCopy_;
  std::vector<Out> outs_;
  std::vector<std::unique_ptr<transport::Buffer>> ins_;
};

}  // namespace glb
