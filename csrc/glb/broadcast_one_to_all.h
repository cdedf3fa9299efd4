// BroadcastOneToAll<T> (old-style): the root sends the whole buffer to every
// peer once each peer has signalled clear-to-send (so back-to-back run()s cannot
// overwrite data that is still being consumed); the result is replicated across
// the local pointers. 1 step, P*S bytes from the root.
// Parity: gloo/broadcast_one_to_all.h:21-102.
#pragma once

#include <cstring>
#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/transport/buffer.h"

namespace glb {

template <typename T>
class BroadcastOneToAll : public Algorithm {
 public:
  BroadcastOneToAll(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs, size_t count,
                    int rootRank = 0, int rootPointerRank = 0)
      : Algorithm(context), ptrs_(ptrs), count_(count), bytes_(count * sizeof(T)), rootRank_(rootRank),
        rootPointerRank_(rootPointerRank) {
    GLB_ENFORCE(rootRank_ >= 0 && rootRank_ < contextSize_, "invalid root rank ", rootRank_);
    GLB_ENFORCE(!ptrs_.empty());
    GLB_ENFORCE(rootPointerRank_ >= 0 && rootPointerRank_ < static_cast<int>(ptrs_.size()));
    if (contextSize_ == 1) return;
    const int dataSlot = context_->nextSlot();
    const int ctsSlot = context_->nextSlot();
    if (contextRank_ == rootRank_) {
      for (int i = 0; i < contextSize_; i++) {
        if (i == rootRank_) continue;
        auto& pair = getPair(i);
        data_.push_back(pair->createSendBuffer(dataSlot, ptrs_[rootPointerRank_], bytes_));
        ctsSink_.push_back(std::make_unique<int>(0));
        cts_.push_back(pair->createRecvBuffer(ctsSlot, ctsSink_.back().get(), sizeof(token_)));
      }
    } else {
      auto& pair = getPair(rootRank_);
      data_.push_back(pair->createRecvBuffer(dataSlot, ptrs_[0], bytes_));
      cts_.push_back(pair->createSendBuffer(ctsSlot, &token_, sizeof(token_)));
    }
  }

  void run() override {
    if (contextSize_ > 1) {
      if (contextRank_ == rootRank_) {
        for (size_t i = 0; i < data_.size(); i++) {
          cts_[i]->waitRecv();
          data_[i]->send();
        }
        for (auto& b : data_) b->waitSend();
      } else {
        cts_[0]->send();
        cts_[0]->waitSend();
        data_[0]->waitRecv();
      }
    }
    // Local fan-out.
    const T* src = contextRank_ == rootRank_ ? ptrs_[rootPointerRank_] : ptrs_[0];
    for (auto* p : ptrs_) {
      if (p != src && bytes_ > 0) std::memcpy(p, src, bytes_);
    }
  }

 protected:
  std::vector<T*> ptrs_;
  const size_t count_;
  const size_t bytes_;
  const int rootRank_;
  const int rootPointerRank_;
  int token_ = 0;  // clear-to-send source; arrivals land in one sink word per peer
  std::vector<std::unique_ptr<int>> ctsSink_;
  std::vector<std::unique_ptr<transport::Buffer>> data_;
  std::vector<std::unique_ptr<transport::Buffer>> cts_;
};

}  // namespace glb
