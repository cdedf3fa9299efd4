// Old-style barriers over bound buffers. Parity: gloo/barrier.h:25-51,
// barrier_all_to_all.h, barrier_all_to_one.h.
#pragma once

#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/transport/buffer.h"

namespace glb {

class Barrier : public Algorithm {
 public:
  explicit Barrier(const std::shared_ptr<Context>& context) : Algorithm(context) {}
  ~Barrier() noexcept(false) override {}
};

// Every rank signals every other rank, then waits for all signals: 1 step, P messages.
class BarrierAllToAll : public Barrier {
 public:
  explicit BarrierAllToAll(const std::shared_ptr<Context>& context) : Barrier(context) {
    const int slot = context_->nextSlot();
    for (int i = 0; i < contextSize_; i++) {
      if (i == contextRank_) continue;
      auto& pair = getPair(i);
      GLB_ENFORCE(pair, "pair missing (rank ", i, ")");
      send_.push_back(pair->createSendBuffer(slot, &token_, sizeof(token_)));
      recv_.push_back(pair->createRecvBuffer(slot, sink(), sizeof(token_)));
    }
  }
  void run() override {
    for (auto& b : send_) b->send();
    for (auto& b : send_) b->waitSend();
    for (auto& b : recv_) b->waitRecv();
  }

 protected:
  // Tokens are sent from token_ (never written) and land in one sink word per peer,
  // so concurrent arrivals on different pairs never touch the same memory.
  int* sink() {
    sinks_.push_back(std::make_unique<int>(0));
    return sinks_.back().get();
  }
  int token_ = 0;
  std::vector<std::unique_ptr<int>> sinks_;
  std::vector<std::unique_ptr<transport::Buffer>> send_;
  std::vector<std::unique_ptr<transport::Buffer>> recv_;
};

// Everyone reports to the root, which then releases everyone: 2 steps.
class BarrierAllToOne : public Barrier {
 public:
  explicit BarrierAllToOne(const std::shared_ptr<Context>& context, int rootRank = 0)
      : Barrier(context), rootRank_(rootRank) {
    GLB_ENFORCE(rootRank >= 0 && rootRank < contextSize_, "invalid root ", rootRank);
    const int slot = context_->nextSlot();
    if (contextRank_ == rootRank_) {
      for (int i = 0; i < contextSize_; i++) {
        if (i == rootRank_) continue;
        auto& pair = getPair(i);
        GLB_ENFORCE(pair, "pair missing (rank ", i, ")");
        send_.push_back(pair->createSendBuffer(slot, &token_, sizeof(token_)));
        recv_.push_back(pair->createRecvBuffer(slot, sink(), sizeof(token_)));
      }
    } else {
      auto& pair = getPair(rootRank_);
      GLB_ENFORCE(pair, "pair missing (rank ", rootRank_, ")");
      send_.push_back(pair->createSendBuffer(slot, &token_, sizeof(token_)));
      recv_.push_back(pair->createRecvBuffer(slot, sink(), sizeof(token_)));
    }
  }
  void run() override {
    if (contextRank_ == rootRank_) {
      for (auto& b : recv_) b->waitRecv();
      for (auto& b : send_) b->send();
      for (auto& b : send_) b->waitSend();
    } else {
      send_[0]->send();
      send_[0]->waitSend();
      recv_[0]->waitRecv();
    }
  }

 protected:
  const int rootRank_;
  // Tokens are sent from token_ (never written) and land in one sink word per peer,
  // so concurrent arrivals on different pairs never touch the same memory.
  int* sink() {
    sinks_.push_back(std::make_unique<int>(0));
    return sinks_.back().get();
  }
  int token_ = 0;
  std::vector<std::unique_ptr<int>> sinks_;
  std::vector<std::unique_ptr<transport::Buffer>> send_;
  std::vector<std::unique_ptr<transport::Buffer>> recv_;
};

}  // namespace glb
