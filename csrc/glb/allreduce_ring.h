// AllreduceRing<T> (old-style): latency-lean ring that forwards the WHOLE buffer
// P-1 times (P*S bytes on the wire): each round a rank sends what it received in
// the previous round to its right neighbour and folds what arrives from the left
// into its result. A consumed-notification keeps a fast sender from overwriting
// an inbox that is still being reduced. Parity: gloo/allreduce_ring.h:20-129.
#pragma once

#include <cstring>
#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/transport/buffer.h"

namespace glb {

template <typename T>
class AllreduceRing : public Algorithm {
 public:
  AllreduceRing(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs, const size_t count,
                const ReductionFunction<T>* fn = ReductionFunction<T>::sum)
      : Algorithm(context), ptrs_(ptrs), count_(count), bytes_(count * sizeof(T)), fn_(fn) {
    GLB_ENFORCE(!ptrs_.empty());
    if (contextSize_ == 1) return;
    inbox_.resize(count_ > 0 ? count_ : 1);
    outbox_.resize(count_ > 0 ? count_ : 1);
    auto& left = getLeftPair();
    auto& right = getRightPair();
    const int dataSlot = context_->nextSlot();
    const int ackSlot = context_->nextSlot();
    sendData_ = right->createSendBuffer(dataSlot, outbox_.data(), bytes_);
    recvData_ = left->createRecvBuffer(dataSlot, inbox_.data(), bytes_);
    sendAck_ = left->createSendBuffer(ackSlot, &token_, sizeof(token_));
    recvAck_ = right->createRecvBuffer(ackSlot, &tokenIn_, sizeof(tokenIn_));
  }

  void run() override {
    if (count_ == 0) return;
    for (size_t i = 1; i < ptrs_.size(); i++) fn_->call(ptrs_[0], ptrs_[i], count_);
    if (contextSize_ > 1) {
      std::memcpy(outbox_.data(), ptrs_[0], bytes_);
      for (int round = 0; round < contextSize_ - 1; round++) {
        sendData_->send();
        recvData_->waitRecv();
        fn_->call(ptrs_[0], inbox_.data(), count_);
        sendData_->waitSend();
        if (round < contextSize_ - 2) std::memcpy(outbox_.data(), inbox_.data(), bytes_);
        // Tell the left neighbour its next write may land; wait for the same from the right.
        sendAck_->send();
        recvAck_->waitRecv();
        sendAck_->waitSend();
      }
    }
    for (size_t i = 1; i < ptrs_.size(); i++) std::memcpy(ptrs_[i], ptrs_[0], bytes_);
  }

 protected:
  std::vector<T*> ptrs_;
  const size_t count_;
  const size_t bytes_;
  const ReductionFunction<T>* fn_;
  std::vector<T> inbox_;
  std::vector<T> outbox_;
  int token_ = 0;    // ack source (never written)
  int tokenIn_ = 0;  // where the neighbour's ack lands
  std::unique_ptr<transport::Buffer> sendData_, recvData_, sendAck_, recvAck_;
};

}  // namespace glb
