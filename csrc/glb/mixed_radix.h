// Shared engine of the old-style recursive algorithms (AllreduceHalvingDoubling,
// AllreduceBcube, ReduceScatterHalvingDoubling): a mixed-radix hypercube over bound
// buffers.
//
// P is written as f_0 * f_1 * ... * f_{k-1}. In reduce-scatter step i a rank splits
// its current block in f_i parts, keeps the part named by its i-th digit, sends the
// others to the f_i - 1 peers that differ from it only in that digit and folds in
// what they send; after k steps every rank owns 1/P of the result. The allgather
// mirrors the steps in reverse, writing straight into the peers' result buffers.
//   factors 2,2,...   => recursive vector halving / distance doubling
//   factors B,B,...   => bcube with base B
// Ranks beyond the largest "regular" size are folded onto partner ranks before and
// after (used for non-power-of-two halving-doubling).
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/collectives_common.h"
#include "glb/common/utils.h"
#include "glb/transport/buffer.h"

namespace glb {
namespace detail {

// Largest divisor of n in [2, base]; n itself when there is none (n prime > base).
inline int pickFactor(int n, int base) {
  for (int f = std::min(n, base); f >= 2; f--) {
    if (n % f == 0) return f;
  }
  return n;
}

inline std::vector<int> radixFactors(int n, int base) {
  std::vector<int> f;
  while (n > 1) {
    int x = pickFactor(n, base);
    f.push_back(x);
    n /= x;
  }
  return f;
}

template <typename T>
class MixedRadix {
 public:
  // `core` ranks [0, coreSize) run the hypercube; rank coreSize + e (an "extra")
  // is folded onto rank e.
  MixedRadix(Algorithm* /*owner*/, const std::shared_ptr<Context>& context, T* data, size_t count,
             const ReductionFunction<T>* fn, std::vector<int> factors, int coreSize, bool allgatherPhase)
      : context_(context), data_(data), count_(count), fn_(fn), factors_(std::move(factors)),
        P_(context->size), r_(context->rank), core_(coreSize), doAllgather_(allgatherPhase) {
    int prod = 1;
    for (int f : factors_) prod *= f;
    GLB_ENFORCE_EQ(prod, core_, "factor list does not multiply to the core size");
    const int K = static_cast<int>(factors_.size());
    const int extras = P_ - core_;
    isExtra_ = r_ >= core_;
    hasExtra_ = !isExtra_ && r_ < extras;
    // fold buffers
    if (extras > 0) {
      const int slotA = context_->nextSlot();
      const int slotB = context_->nextSlot();
      if (isExtra_) {
        auto& pair = context_->getPair(r_ - core_);
        foldSend_ = pair->createSendBuffer(slotA, data_, count_ * sizeof(T));
        foldRecv_ = pair->createRecvBuffer(slotB, data_, count_ * sizeof(T));
      } else if (hasExtra_) {
        foldScratch_.resize(std::max<size_t>(count_, 1));
        auto& pair = context_->getPair(r_ + core_);
        foldRecv_ = pair->createRecvBuffer(slotA, foldScratch_.data(), count_ * sizeof(T));
        foldSend_ = pair->createSendBuffer(slotB, data_, count_ * sizeof(T));
      }
    }
    // Every rank reserves the same slots, whether it uses them or not, so later
    // algorithm instances on this context stay in step.
    const int rsSlot = context_->nextSlot(std::max(K, 1));
    const int agSlot = context_->nextSlot(std::max(K, 1));
    const int ackSlot = context_->nextSlot(std::max(K, 1));
    if (isExtra_) return;

    stride_.resize(K);
    digit_.resize(K);
    int s = 1;
    for (int i = 0; i < K; i++) {
      stride_[i] = s;
      digit_[i] = (r_ / s) % factors_[i];
      s *= factors_[i];
    }
    blocks_.resize(K + 1);
    blocks_[0] = Range{0, count_};
    for (int i = 0; i < K; i++) blocks_[i + 1] = subRange(blocks_[i], factors_[i], digit_[i]);

    steps_.resize(K);
    for (int i = 0; i < K; i++) {
      auto& st = steps_[i];
      const int f = factors_[i];
      const Range mine = blocks_[i + 1];
      st.scratch.resize(std::max<size_t>(1, static_cast<size_t>(f - 1) * mine.len));
      int k = 0;
      for (int d = 0; d < f; d++) {
        if (d == digit_[i]) continue;
        const int peer = r_ + (d - digit_[i]) * stride_[i];
        auto& pair = context_->getPair(peer);
        GLB_ENFORCE(pair, "pair missing (rank ", peer, ")");
        Peer p;
        p.theirs = subRange(blocks_[i], f, d);
        p.scratchOff = static_cast<size_t>(k) * mine.len;
        p.rsSend = pair->createSendBuffer(rsSlot + i, data_, count_ * sizeof(T));
        p.rsRecv = pair->createRecvBuffer(rsSlot + i, st.scratch.data() + p.scratchOff, mine.len * sizeof(T));
        if (doAllgather_) {
          p.agSend = pair->createSendBuffer(agSlot + i, data_, count_ * sizeof(T));
          p.agRecv = pair->createRecvBuffer(agSlot + i, data_, count_ * sizeof(T));
        } else {
          // Without the allgather phase nothing flows back to the sender, so a fast
          // peer could start its next run and overwrite this rank's landing zone
          // while it is still being reduced. One credit per (step, peer): the
          // receiver returns it after consuming, the sender needs it to send again.
          p.creditIn = std::make_unique<int>(0);
          p.ackSend = pair->createSendBuffer(ackSlot + i, &credit_, sizeof(credit_));
          p.ackRecv = pair->createRecvBuffer(ackSlot + i, p.creditIn.get(), sizeof(credit_));
        }
        st.peers.push_back(std::move(p));
        k++;
      }
    }
  }

  // Block this rank owns after the reduce-scatter phase (empty for folded-away ranks).
  Range ownedBlock() const { return isExtra_ ? Range{0, 0} : blocks_.back(); }
  // Block rank `rank` owns (for redistribution).
  Range ownedBlockOf(int rank) const {
    if (rank >= core_) return Range{0, 0};
    Range b{0, count_};
    int s = 1;
    for (size_t i = 0; i < factors_.size(); i++) {
      b = subRange(b, factors_[i], (rank / s) % factors_[i]);
      s *= factors_[i];
    }
    return b;
  }
  bool isExtra() const { return isExtra_; }

  void foldIn() {
    if (isExtra_) {
      foldSend_->send();
      foldSend_->waitSend();
    } else if (hasExtra_) {
      foldRecv_->waitRecv();
      fn_->call(data_, foldScratch_.data(), count_);
    }
  }

  void foldOut(size_t offsetElems, size_t lenElems) {
    if (isExtra_) {
      foldRecv_->waitRecv();
    } else if (hasExtra_) {
      foldSend_->send(offsetElems * sizeof(T), lenElems * sizeof(T), offsetElems * sizeof(T));
      foldSend_->waitSend();
    }
  }

  void reduceScatter() {
    if (isExtra_) return;
    const int K = static_cast<int>(factors_.size());
    for (int i = 0; i < K; i++) {
      auto& st = steps_[i];
      const Range mine = blocks_[i + 1];
      for (auto& p : st.peers) {
        if (!doAllgather_ && runs_ > 0) p.ackRecv->waitRecv();
        p.rsSend->send(p.theirs.off * sizeof(T), p.theirs.len * sizeof(T), 0);
      }
      for (auto& p : st.peers) {
        p.rsRecv->waitRecv();
        if (mine.len > 0) fn_->call(data_ + mine.off, st.scratch.data() + p.scratchOff, mine.len);
        if (!doAllgather_) p.ackSend->send();
      }
      for (auto& p : st.peers) {
        p.rsSend->waitSend();
        if (!doAllgather_) p.ackSend->waitSend();
      }
    }
    runs_++;
  }

  void allgather() {
    if (isExtra_) return;
    const int K = static_cast<int>(factors_.size());
    for (int i = K - 1; i >= 0; i--) {
      auto& st = steps_[i];
      const Range mine = blocks_[i + 1];
      for (auto& p : st.peers) p.agSend->send(mine.off * sizeof(T), mine.len * sizeof(T), mine.off * sizeof(T));
      for (auto& p : st.peers) p.agRecv->waitRecv();
      for (auto& p : st.peers) p.agSend->waitSend();
    }
  }

  int numSteps() const { return static_cast<int>(factors_.size()); }

 private:
  struct Peer {
    Range theirs;
    size_t scratchOff = 0;
    // Declared before the buffers: members die in reverse order, and the landing
    // word must outlive the recv buffer that points at it.
    std::unique_ptr<int> creditIn;
    std::unique_ptr<transport::Buffer> rsSend, rsRecv, agSend, agRecv, ackSend, ackRecv;
  };
  struct Step {
    std::vector<T> scratch;
    std::vector<Peer> peers;
  };

  std::shared_ptr<Context> context_;
  T* data_;
  const size_t count_;
  const ReductionFunction<T>* fn_;
  std::vector<int> factors_;
  const int P_;
  const int r_;
  const int core_;
  const bool doAllgather_;
  int credit_ = 0;     // credit source (never written)
  uint64_t runs_ = 0;
  bool isExtra_ = false;
  bool hasExtra_ = false;
  std::vector<int> stride_, digit_;
  std::vector<Range> blocks_;
  std::vector<Step> steps_;
  std::vector<T> foldScratch_;
  std::unique_ptr<transport::Buffer> foldSend_, foldRecv_;
};

inline int largestPow2AtMost(int n) {
  int p = 1;
  while (p * 2 <= n) p *= 2;
  return p;
}

}  // namespace detail
}  // namespace glb
