// PairwiseExchange (benchmark communication pattern): each rank exchanges
// numBytes / numDestinations bytes with its XOR partners rank ^ 2^k, one partner per
// step. Parity: gloo/pairwise_exchange.h:19-73.
#pragma once

#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/common/utils.h"
#include "glb/transport/buffer.h"

namespace glb {

class PairwiseExchange : public Algorithm {
 public:
  PairwiseExchange(const std::shared_ptr<Context>& context, const int numBytes, const int numDestinations)
      : Algorithm(context), numDestinations_(numDestinations),
        bytesPerMsg_(numDestinations > 0 ? numBytes / numDestinations : 0),
        sendBuf_(std::max(1, numBytes)), recvBuf_(std::max(1, numBytes)) {
    GLB_ENFORCE(isPow2(static_cast<uint64_t>(contextSize_)), "pairwise_exchange needs a power-of-two context size");
    GLB_ENFORCE_GT(numDestinations_, 0);
    GLB_ENFORCE_LE(numDestinations_, static_cast<int>(log2ceil(static_cast<uint32_t>(contextSize_))),
                   "at most log2(P) destinations");
    int bit = 1;
    for (int i = 0; i < numDestinations_; i++, bit <<= 1) {
      const int peer = contextRank_ ^ bit;
      auto& pair = getPair(peer);
      const int slot = context_->nextSlot();
      send_.push_back(pair->createSendBuffer(slot, sendBuf_.data() + i * bytesPerMsg_, bytesPerMsg_));
      recv_.push_back(pair->createRecvBuffer(slot, recvBuf_.data() + i * bytesPerMsg_, bytesPerMsg_));
    }
  }

  void run() override {
    for (int i = 0; i < numDestinations_; i++) {
      send_[i]->send();
      recv_[i]->waitRecv();
      send_[i]->waitSend();
    }
  }

 protected:
  const int numDestinations_;
  const int bytesPerMsg_;
  std::vector<char> sendBuf_, recvBuf_;
  std::vector<std::unique_ptr<transport::Buffer>> send_, recv_;
};

}  // namespace glb
