// Bound buffer on a tcp::Pair: completion counters + condvars (async mode) or
// caller-driven progress (sync mode). Parity: gloo/transport/tcp/buffer.{h,cc}.
#pragma once

#include <condition_variable>
#include <exception>
#include <mutex>
#include <string>

#include "glb/transport/buffer.h"

namespace glb {
namespace transport {
namespace tcp {

class Pair;

class Buffer : public ::glb::transport::Buffer {
 public:
  Buffer(Pair* pair, int slot, void* ptr, size_t size, bool isRecv);
  ~Buffer() override;

  void send(size_t offset, size_t length, size_t roffset = 0) override;
  using ::glb::transport::Buffer::send;
  void waitRecv() override;
  void waitSend() override;

  // Called by the pair with its mutex held.
  void handleRecvCompletion();
  void handleSendCompletion();
  void signalException(const std::string& msg);

 private:
  void throwIfException();  // requires pair mutex

  Pair* pair_;
  const bool isRecv_;
  std::condition_variable recvCv_;
  std::condition_variable sendCv_;
  int recvCompletions_ = 0;
  int sendCompletions_ = 0;
  int sendPending_ = 0;
  bool failed_ = false;
  std::string exMsg_;
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
