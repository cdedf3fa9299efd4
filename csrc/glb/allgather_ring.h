// AllgatherRing<T> (old-style): several input pointers per rank; the output holds
// P * inputs blocks, block (rank, i) at (rank * inputs + i) * count. P-1 rounds per
// input; each round forwards the block received in the previous round, written
// straight into the right neighbour's output with send(offset, len, roffset).
// Parity: gloo/allgather_ring.h:26-111.
#pragma once

#include <cstring>
#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/transport/buffer.h"

namespace glb {

template <typename T>
class AllgatherRing : public Algorithm {
 public:
  AllgatherRing(const std::shared_ptr<Context>& context, const std::vector<const T*>& inPtrs, T* outPtr,
                size_t count)
      : Algorithm(context), inPtrs_(inPtrs), outPtr_(outPtr), count_(count), bytes_(count * sizeof(T)),
        inputStride_(count * inPtrs.size()) {
    GLB_ENFORCE(!inPtrs_.empty());
    GLB_ENFORCE(outPtr_ != nullptr);
    if (contextSize_ == 1) return;
    auto& left = getLeftPair();
    auto& right = getRightPair();
    const size_t total = inPtrs_.size() * contextSize_ * bytes_;
    const int dataSlot = context_->nextSlot();
    const int ackSlot = context_->nextSlot();
    sendData_ = right->createSendBuffer(dataSlot, outPtr_, total);
    recvData_ = left->createRecvBuffer(dataSlot, outPtr_, total);
    sendAck_ = left->createSendBuffer(ackSlot, &token_, sizeof(token_));
    recvAck_ = right->createRecvBuffer(ackSlot, &tokenIn_, sizeof(tokenIn_));
  }

  void run() override {
    const int P = contextSize_;
    const int r = contextRank_;
    for (size_t i = 0; i < inPtrs_.size(); i++) {
      if (bytes_ > 0) std::memcpy(outPtr_ + r * inputStride_ + i * count_, inPtrs_[i], bytes_);
    }
    if (P == 1) return;
    const size_t blockBytes = inputStride_ * sizeof(T);
    for (int round = 0; round < P - 1; round++) {
      const int sendRank = (r - round + P) % P;
      for (size_t i = 0; i < inPtrs_.size(); i++) {
        const size_t off = sendRank * blockBytes + i * bytes_;
        sendData_->send(off, bytes_, off);
      }
      for (size_t i = 0; i < inPtrs_.size(); i++) recvData_->waitRecv();
      for (size_t i = 0; i < inPtrs_.size(); i++) sendData_->waitSend();
    }
    // Neighbour handshake so a following run() cannot race with this one's tail.
    sendAck_->send();
    recvAck_->waitRecv();
    sendAck_->waitSend();
  }

 protected:
  std::vector<const T*> inPtrs_;
  T* outPtr_;
  const size_t count_;
  const size_t bytes_;
  const size_t inputStride_;
  int token_ = 0;    // ack source (never written)
  int tokenIn_ = 0;  // where the neighbour's ack lands
  std::unique_ptr<transport::Buffer> sendData_, recvData_, sendAck_, recvAck_;
};

}  // namespace glb
