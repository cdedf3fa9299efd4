// tcp::UnboundBuffer. Parity: gloo/transport/tcp/unbound_buffer.{h,cc}.
#pragma once

#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <vector>

#include "glb/common/memory.h"
#include "glb/transport/unbound_buffer.h"

namespace glb {
namespace transport {
namespace tcp {

class Context;

class UnboundBuffer : public ::glb::transport::UnboundBuffer {
 public:
  UnboundBuffer(std::shared_ptr<Context> context, void* ptr, size_t size);
  ~UnboundBuffer() override;

  bool waitRecv(int* rank, std::chrono::milliseconds timeout) override;
  bool waitSend(int* rank, std::chrono::milliseconds timeout) override;
  using ::glb::transport::UnboundBuffer::waitRecv;
  using ::glb::transport::UnboundBuffer::waitSend;
  void abortWaitRecv() override;
  void abortWaitSend() override;

  void send(int dstRank, uint64_t slot, size_t offset, size_t nbytes) override;
  void recv(int srcRank, uint64_t slot, size_t offset, size_t nbytes) override;
  void recv(std::vector<int> srcRanks, uint64_t slot, size_t offset, size_t nbytes) override;

  std::unique_ptr<::glb::transport::RemoteKey> getRemoteKey() const override;
  void put(const ::glb::transport::RemoteKey& key, uint64_t slot, size_t offset, size_t roffset,
           size_t nbytes) override;
  void get(const ::glb::transport::RemoteKey& key, uint64_t slot, size_t offset, size_t roffset,
           size_t nbytes) override;

  // Completion hooks (I/O thread or the posting thread on immediate completion).
  void handleRecvCompletion(int rank);
  void handleSendCompletion(int rank);
  void signalException(const std::string& msg);

  WeakAnchor<UnboundBuffer> weak() const { return anchor_.weak(); }

 private:
  void throwIfException();  // requires m_
  void notePeer(int rank);  // requires no lock
  // Spin-then-block (see Pair::spinWait): poll the pairs this buffer talks to for a
  // bounded time or until done() holds. `lock` holds m_ on entry and on return.
  template <typename Pred>
  void spinUntil(std::unique_lock<std::mutex>& lock, Pred done);
  // Pairs in sync mode are detached from the loop thread, so the waiter has to read
  // them itself for as long as it waits (the reference rejects unbound buffers on
  // sync pairs outright). Returns false on timeout; true when done() holds or when
  // no peer of this buffer is in sync mode (then the caller blocks on the condvar).
  template <typename Pred>
  bool driveSyncPairs(std::unique_lock<std::mutex>& lock, Pred done, std::chrono::milliseconds timeout);

  std::shared_ptr<Context> context_;
  std::mutex m_;
  std::condition_variable recvCv_;
  std::condition_variable sendCv_;
  bool abortWaitRecv_ = false;
  bool abortWaitSend_ = false;
  std::deque<int> recvRanks_;  // one entry per completed recv
  std::deque<int> sendRanks_;  // one entry per completed send
  std::vector<int> spinRanks_;  // peers this buffer ever sent to / received from (sticky, small)
  bool failed_ = false;
  std::string exMsg_;
  mutable uint64_t regionId_ = 0;
  Anchor<UnboundBuffer> anchor_;
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
