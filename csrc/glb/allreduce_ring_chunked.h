// AllreduceRingChunked<T> (old-style): bandwidth-optimal chunked ring. The buffer
// is cut into 2P chunks handled as two interleaved lanes (even / odd chunks) so
// that one lane's transfer overlaps the other lane's reduction. Per lane: P-1
// reduce-scatter rounds through a small inbox, then P-1 allgather rounds written
// straight into the right neighbour's result. ~4P steps, 2·S·(P-1)/P bytes.
// Flow control: a rank may refill its neighbour's inbox only after the
// neighbour acknowledged the previous content (per-lane ack buffers).
// Parity: gloo/allreduce_ring_chunked.h:20-254.
#pragma once

#include <cstring>
#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/collectives_common.h"
#include "glb/common/utils.h"
#include "glb/transport/buffer.h"

namespace glb {

template <typename T>
class AllreduceRingChunked : public Algorithm {
 public:
  AllreduceRingChunked(const std::shared_ptr<Context>& context, const std::vector<T*>& ptrs, const size_t count,
                       const ReductionFunction<T>* fn = ReductionFunction<T>::sum)
      : Algorithm(context), ptrs_(ptrs), count_(count), bytes_(count * sizeof(T)), fn_(fn) {
    GLB_ENFORCE(!ptrs_.empty());
    if (contextSize_ == 1) return;
    const int P = contextSize_;
    // 2P near-equal chunks; chunk c belongs to lane c & 1, ring position c >> 1.
    maxChunk_ = ceilDiv(std::max<size_t>(count_, 1), static_cast<size_t>(2 * P));
    auto& left = getLeftPair();
    auto& right = getRightPair();
    for (int l = 0; l < 2; l++) {
      inbox_[l].resize(maxChunk_);
      const int slot = context_->nextSlot();
      const int ackSlot = context_->nextSlot();
      const int agSlot = context_->nextSlot();
      rsSend_[l] = right->createSendBuffer(slot, ptrs_[0], bytes_);
      rsRecv_[l] = left->createRecvBuffer(slot, inbox_[l].data(), maxChunk_ * sizeof(T));
      ackSend_[l] = left->createSendBuffer(ackSlot, &token_, sizeof(token_));
      ackRecv_[l] = right->createRecvBuffer(ackSlot, &tokenIn_[l], sizeof(token_));
      agSend_[l] = right->createSendBuffer(agSlot, ptrs_[0], bytes_);
      agRecv_[l] = left->createRecvBuffer(agSlot, ptrs_[0], bytes_);
    }
  }

  void run() override {
    if (count_ == 0) return;
    for (size_t i = 1; i < ptrs_.size(); i++) fn_->call(ptrs_[0], ptrs_[i], count_);
    if (contextSize_ > 1) {
      const int P = contextSize_;
      const int r = contextRank_;
      T* data = ptrs_[0];
      auto chunk = [&](int lane, int pos) {  // pos in [0, P)
        return detail::subRange(detail::Range{0, count_}, 2 * P, 2 * ((pos % P + P) % P) + lane);
      };
      // Reduce-scatter: in round s lane l sends chunk (r - s) and receives chunk (r - s - 1).
      for (int s = 0; s < P - 1; s++) {
        for (int l = 0; l < 2; l++) {
          if (s > 0) ackRecv_[l]->waitRecv();  // right neighbour emptied its inbox
          auto c = chunk(l, r - s);
          rsSend_[l]->send(c.off * sizeof(T), c.len * sizeof(T), 0);
        }
        for (int l = 0; l < 2; l++) {
          rsRecv_[l]->waitRecv();
          auto c = chunk(l, r - s - 1);
          if (c.len > 0) fn_->call(data + c.off, inbox_[l].data(), c.len);
          if (s < P - 2) ackSend_[l]->send();
        }
        for (int l = 0; l < 2; l++) {
          rsSend_[l]->waitSend();
          if (s < P - 2) ackSend_[l]->waitSend();
        }
      }
      // Allgather: rank r owns chunk (r + 1); round s forwards chunk (r + 1 - s).
      for (int s = 0; s < P - 1; s++) {
        for (int l = 0; l < 2; l++) {
          auto c = chunk(l, r + 1 - s);
          agSend_[l]->send(c.off * sizeof(T), c.len * sizeof(T), c.off * sizeof(T));
        }
        for (int l = 0; l < 2; l++) agRecv_[l]->waitRecv();
        for (int l = 0; l < 2; l++) agSend_[l]->waitSend();
      }
      // Closing handshake: nobody starts the next run (and overwrites a neighbour's
      // inbox or result) before its neighbours are done with this one.
      for (int l = 0; l < 2; l++) ackSend_[l]->send();
      for (int l = 0; l < 2; l++) ackRecv_[l]->waitRecv();
      for (int l = 0; l < 2; l++) ackSend_[l]->waitSend();
    }
    for (size_t i = 1; i < ptrs_.size(); i++) std::memcpy(ptrs_[i], ptrs_[0], bytes_);
  }

 protected:
  std::vector<T*> ptrs_;
  const size_t count_;
  const size_t bytes_;
  const ReductionFunction<T>* fn_;
  size_t maxChunk_ = 0;
  int token_ = 0;             // ack source (never written)
  int tokenIn_[2] = {0, 0};   // where the right neighbour's acks land, per lane
  std::vector<T> inbox_[2];
  std::unique_ptr<transport::Buffer> rsSend_[2], rsRecv_[2], ackSend_[2], ackRecv_[2], agSend_[2], agRecv_[2];
};

}  // namespace glb
