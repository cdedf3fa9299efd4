// broadcast (new-style): binomial tree rooted at `root` (virtual rank
// (rank - root) mod P), ceil(log2 P) rounds. The root may pass a separate input.
// Large payloads are pipelined down the tree in segments so interior nodes forward
// while still receiving. Parity: gloo/broadcast.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class BroadcastOptions : public detail::CollectiveOptionsBase {
 public:
  explicit BroadcastOptions(const std::shared_ptr<Context>& context) : CollectiveOptionsBase(context) {}

  template <typename T>
  void setInput(std::unique_ptr<UnboundBuffer> buf) {
    elementSize = sizeof(T);
    in = std::move(buf);
  }
  template <typename T>
  void setInput(T* ptr, size_t n) {
    elementSize = sizeof(T);
    in = context->createUnboundBuffer(ptr, n * sizeof(T));
  }
  template <typename T>
  void setOutput(std::unique_ptr<UnboundBuffer> buf) {
    elementSize = sizeof(T);
    out = std::move(buf);
  }
  template <typename T>
  void setOutput(T* ptr, size_t n) {
    elementSize = sizeof(T);
    out = context->createUnboundBuffer(ptr, n * sizeof(T));
  }
  void setInputRaw(void* ptr, size_t bytes) { in = context->createUnboundBuffer(ptr, bytes); elementSize = 1; }
  void setOutputRaw(void* ptr, size_t bytes) { out = context->createUnboundBuffer(ptr, bytes); elementSize = 1; }
  void setRoot(int r) { root = r; }
  void setMaxSegmentSize(size_t s) { maxSegmentSize = s; }

  std::unique_ptr<UnboundBuffer> in;   // optional, root only
  std::unique_ptr<UnboundBuffer> out;  // required on every rank
  size_t elementSize = 0;
  int root = -1;
  size_t maxSegmentSize = 4u << 20;
};

void broadcast(BroadcastOptions& opts);

}  // namespace glb
