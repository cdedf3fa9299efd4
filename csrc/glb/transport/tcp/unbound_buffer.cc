#include "glb/transport/tcp/unbound_buffer.h"

#include <sched.h>

#include <memory>
#include <vector>

#include "glb/common/logging.h"
#include "glb/transport/tcp/context.h"
#include "glb/transport/tcp/pair.h"

namespace glb {
namespace transport {
namespace tcp {

UnboundBuffer::UnboundBuffer(std::shared_ptr<Context> context, void* ptr, size_t size)
    : ::glb::transport::UnboundBuffer(ptr, size), context_(std::move(context)), anchor_(this) {}

UnboundBuffer::~UnboundBuffer() {
  // Remove every reference the transport may still hold before the memory goes away.
  context_->forgetBuffer(this, regionId_);
  anchor_.retire();
}

void UnboundBuffer::throwIfException() {
  if (failed_) GLB_THROW_IO_EXCEPTION(exMsg_);
}

void UnboundBuffer::handleRecvCompletion(int rank) {
  std::lock_guard<std::mutex> g(m_);
  recvRanks_.push_back(rank);
  recvCv_.notify_one();
}

void UnboundBuffer::handleSendCompletion(int rank) {
  std::lock_guard<std::mutex> g(m_);
  sendRanks_.push_back(rank);
  sendCv_.notify_one();
}

void UnboundBuffer::signalException(const std::string& msg) {
  std::lock_guard<std::mutex> g(m_);
  failed_ = true;
  exMsg_ = msg;
  recvCv_.notify_all();
  sendCv_.notify_all();
}

void UnboundBuffer::notePeer(int rank) {
  std::lock_guard<std::mutex> g(m_);
  for (int r : spinRanks_) {
    if (r == rank) return;
  }
  spinRanks_.push_back(rank);
}

template <typename Pred>
void UnboundBuffer::spinUntil(std::unique_lock<std::mutex>& lock, Pred done) {
  const int64_t budget = Pair::spinBudgetNanos();
  if (budget == 0 || spinRanks_.empty()) return;
  const std::vector<int> ranks = spinRanks_;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::nanoseconds(budget);
  unsigned spins = 0;
  // Ring-like patterns talk to one or two peers: park their descriptors so the loop thread
  // sleeps through the arrivals this thread is about to consume (two syscalls per peer;
  // not worth it for wide fan-in such as alltoall).
  std::vector<std::unique_ptr<Pair::PauseGuard>> quiet;
  if (ranks.size() <= 2 && !done()) {
    lock.unlock();
    for (int r : ranks) {
      auto* p = static_cast<Pair*>(context_->peekPair(r));
      if (p != nullptr) quiet.push_back(std::make_unique<Pair::PauseGuard>(p, /*locked=*/false));
    }
    lock.lock();
  }
  struct Unpark {  // guards take pair mutexes: release them without holding m_
    std::unique_lock<std::mutex>& lock;
    std::vector<std::unique_ptr<Pair::PauseGuard>>& q;
    ~Unpark() {
      if (q.empty()) return;
      lock.unlock();
      q.clear();
      lock.lock();
    }
  } unpark{lock, quiet};
  while (!done()) {
    // Pair mutex before m_ is the completion path's order, so m_ is dropped here.
    lock.unlock();
    for (int r : ranks) {
      auto* p = static_cast<Pair*>(context_->peekPair(r));
      if (p != nullptr) p->tryProgress();
    }
    if ((++spins & 31) == 0) ::sched_yield();  // see Pair::spinWait
    lock.lock();
    if (std::chrono::steady_clock::now() >= deadline) return;
  }
}

template <typename Pred>
bool UnboundBuffer::driveSyncPairs(std::unique_lock<std::mutex>& lock, Pred done,
                                   std::chrono::milliseconds timeout) {
  // Every sync pair of the context, not only this buffer's peers: with no loop thread,
  // whatever arrives anywhere (acknowledgements for another buffer's sends, messages that
  // belong in the unexpected queue) is only ever read by a waiting thread.
  std::vector<Pair*> pairs;
  bool busy = true;
  for (int r = 0; r < context_->size; r++) {
    if (r == context_->rank) continue;
    auto* p = static_cast<Pair*>(context_->peekPair(r));
    if (p != nullptr && p->isSync()) {
      pairs.push_back(p);
      busy = busy && p->isBusyPoll();
    }
  }
  if (pairs.empty()) return true;
  const auto start = std::chrono::steady_clock::now();
  while (!done()) {
    lock.unlock();
    for (auto* p : pairs) p->tryProgress();
    lock.lock();
    if (done()) break;
    if (timeout != kNoTimeout && std::chrono::steady_clock::now() - start > timeout) return false;
    // Completions from async pairs still arrive through the condvar.
    if (!busy) recvCv_.wait_for(lock, std::chrono::microseconds(200));
  }
  return true;
}

bool UnboundBuffer::waitRecv(int* rank, std::chrono::milliseconds timeout) {
  if (timeout == kUnsetTimeout) timeout = context_->getTimeout();
  std::unique_lock<std::mutex> lock(m_);
  // Completions that landed before a failure are still handed out.
  if (recvRanks_.empty()) throwIfException();
  if (recvRanks_.empty()) spinUntil(lock, [&] { return abortWaitRecv_ || !recvRanks_.empty() || failed_; });
  if (recvRanks_.empty()) {
    auto pred = [&] { return abortWaitRecv_ || !recvRanks_.empty() || failed_; };
    bool done = driveSyncPairs(lock, pred, timeout);
    if (!done) {
    } else if (timeout == kNoTimeout) {
      recvCv_.wait(lock, pred);
    } else {
      done = recvCv_.wait_for(lock, timeout, pred);
    }
    if (!done) {
      lock.unlock();
      // A stuck receive leaves the wire in an unknown state for every peer that
      // may be involved, so the whole context is poisoned (reference behaviour:
      // tcp/unbound_buffer.cc:52-94).
      context_->signalException("Application timeout caused pair closure");
      GLB_THROW_TIMEOUT("Timed out waiting ", timeout.count(), "ms for recv operation to complete");
    }
    if (recvRanks_.empty()) throwIfException();
  }
  if (abortWaitRecv_ && recvRanks_.empty()) {
    abortWaitRecv_ = false;
    return false;
  }
  abortWaitRecv_ = false;
  if (rank != nullptr) *rank = recvRanks_.front();
  recvRanks_.pop_front();
  return true;
}

bool UnboundBuffer::waitSend(int* rank, std::chrono::milliseconds timeout) {
  if (timeout == kUnsetTimeout) timeout = context_->getTimeout();
  std::unique_lock<std::mutex> lock(m_);
  throwIfException();
  if (sendRanks_.empty()) spinUntil(lock, [&] { return abortWaitSend_ || !sendRanks_.empty() || failed_; });
  if (sendRanks_.empty()) {
    auto pred = [&] { return abortWaitSend_ || !sendRanks_.empty() || failed_; };
    bool done = driveSyncPairs(lock, pred, timeout);
    if (!done) {
    } else if (timeout == kNoTimeout) {
      sendCv_.wait(lock, pred);
    } else {
      done = sendCv_.wait_for(lock, timeout, pred);
    }
    if (!done) {
      lock.unlock();
      context_->signalException("Application timeout caused pair closure");
      GLB_THROW_TIMEOUT("Timed out waiting ", timeout.count(), "ms for send operation to complete");
    }
    throwIfException();
  }
  if (abortWaitSend_ && sendRanks_.empty()) {
    abortWaitSend_ = false;
    return false;
  }
  abortWaitSend_ = false;
  if (rank != nullptr) *rank = sendRanks_.front();
  sendRanks_.pop_front();
  return true;
}

void UnboundBuffer::abortWaitRecv() {
  context_->cancelPostedRecvs(this);
  std::lock_guard<std::mutex> g(m_);
  abortWaitRecv_ = true;
  recvCv_.notify_one();
}

void UnboundBuffer::abortWaitSend() {
  std::lock_guard<std::mutex> g(m_);
  abortWaitSend_ = true;
  sendCv_.notify_one();
}

void UnboundBuffer::send(int dstRank, uint64_t slot, size_t offset, size_t nbytes) {
  if (nbytes == kUnspecifiedByteCount) {
    GLB_ENFORCE_LE(offset, size);
    nbytes = size - offset;
  }
  GLB_ENFORCE_LE(offset + nbytes, size, "send range exceeds buffer");
  notePeer(dstRank);
  context_->tcpPair(dstRank)->sendUnbound(this, slot, offset, nbytes);
}

void UnboundBuffer::recv(int srcRank, uint64_t slot, size_t offset, size_t nbytes) {
  recv(std::vector<int>{srcRank}, slot, offset, nbytes);
}

void UnboundBuffer::recv(std::vector<int> srcRanks, uint64_t slot, size_t offset, size_t nbytes) {
  if (nbytes == kUnspecifiedByteCount) {
    GLB_ENFORCE_LE(offset, size);
    nbytes = size - offset;
  }
  GLB_ENFORCE_LE(offset + nbytes, size, "recv range exceeds buffer");
  for (int r : srcRanks) notePeer(r);
  context_->postRecv(this, std::move(srcRanks), slot, offset, nbytes);
}

std::unique_ptr<::glb::transport::RemoteKey> UnboundBuffer::getRemoteKey() const {
  if (regionId_ == 0) regionId_ = context_->registerRegion(const_cast<UnboundBuffer*>(this));
  return std::make_unique<RemoteKey>(context_->rank, size, regionId_);
}

void UnboundBuffer::put(const ::glb::transport::RemoteKey& key, uint64_t /*slot*/, size_t offset,
                        size_t roffset, size_t nbytes) {
  auto* k = dynamic_cast<const RemoteKey*>(&key);
  GLB_ENFORCE(k != nullptr, "put: not a tcp RemoteKey");
  GLB_ENFORCE_LE(offset + nbytes, size, "put: local range exceeds buffer");
  GLB_ENFORCE_LE(roffset + nbytes, k->size, "put: remote range exceeds region");
  GLB_ENFORCE_NE(k->rank, context_->rank, "put to self");
  notePeer(k->rank);
  context_->tcpPair(k->rank)->sendPut(this, k->regionId, offset, roffset, nbytes);
}

void UnboundBuffer::get(const ::glb::transport::RemoteKey& key, uint64_t /*slot*/, size_t offset,
                        size_t roffset, size_t nbytes) {
  auto* k = dynamic_cast<const RemoteKey*>(&key);
  GLB_ENFORCE(k != nullptr, "get: not a tcp RemoteKey");
  GLB_ENFORCE_LE(offset + nbytes, size, "get: local range exceeds buffer");
  GLB_ENFORCE_LE(roffset + nbytes, k->size, "get: remote range exceeds region");
  GLB_ENFORCE_NE(k->rank, context_->rank, "get from self");
  notePeer(k->rank);
  uint64_t req = context_->registerPendingGet(this, offset, nbytes);
  context_->tcpPair(k->rank)->sendGetRequest(req, k->regionId, roffset, nbytes);
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
