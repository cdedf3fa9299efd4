#include "glb/transport/tcp/buffer.h"

#include <cstdio>
#include <unistd.h>

#include "glb/common/logging.h"
#include "glb/transport/tcp/pair.h"

namespace glb {
namespace transport {
namespace tcp {

// The pair mutex guards the counters: completions are produced by the pair while
// it holds that mutex, so one lock serves both and there is no lock-order issue.
Buffer::Buffer(Pair* pair, int slot, void* ptr, size_t size, bool isRecv)
    : ::glb::transport::Buffer(slot, ptr, size), pair_(pair), isRecv_(isRecv) {}

Buffer::~Buffer() { pair_->unregisterBuffer(this); }

void Buffer::throwIfException() {
  if (failed_) GLB_THROW_IO_EXCEPTION(exMsg_);
}

void Buffer::handleRecvCompletion() {
  recvCompletions_++;
  recvCv_.notify_one();
}

void Buffer::handleSendCompletion() {
  sendCompletions_++;
  sendPending_--;
  sendCv_.notify_one();
}

void Buffer::signalException(const std::string& msg) {
  failed_ = true;
  exMsg_ = msg;
  recvCv_.notify_all();
  sendCv_.notify_all();
}

void Buffer::send(size_t offset, size_t length, size_t roffset) {
  if (debug_) {
    std::fprintf(stderr, "[glb pid %d] tcp::Buffer::send slot=%d offset=%zu length=%zu roffset=%zu\n",
                 static_cast<int>(::getpid()), slot_, offset, length, roffset);
  }
  GLB_ENFORCE_LE(offset + length, size_, "send range exceeds buffer");
  {
    std::lock_guard<std::mutex> g(pair_->mu());
    throwIfException();
    sendPending_++;
  }
  pair_->sendBound(this, offset, length, roffset);
}

void Buffer::waitRecv() {
  std::unique_lock<std::mutex> lock(pair_->mu());
  auto pred = [&] { return recvCompletions_ > 0 || failed_; };
  if (pair_->isSync()) {
    pair_->syncWait(lock, pred, pair_->timeout(), "recv");
  } else {
    pair_->spinWait(lock, pred);
    auto timeout = pair_->timeout();
    if (timeout == kNoTimeout) {
      recvCv_.wait(lock, pred);
    } else if (!recvCv_.wait_for(lock, timeout, pred)) {
      lock.unlock();
      auto msg = strcat_all("Timed out waiting ", timeout.count(), "ms for recv operation to complete");
      pair_->signalExceptionExternal(msg);
      GLB_THROW_TIMEOUT(msg);
    }
  }
  // A write that landed before the pair failed is still delivered.
  if (recvCompletions_ == 0) throwIfException();
  recvCompletions_--;
}

void Buffer::waitSend() {
  std::unique_lock<std::mutex> lock(pair_->mu());
  // waitSend consumes one completion if a send is outstanding or already done;
  // with nothing pending it returns immediately (matches reference semantics where
  // waitSend after a synchronous write is a no-op).
  auto pred = [&] { return sendCompletions_ > 0 || sendPending_ == 0 || failed_; };
  if (pair_->isSync()) {
    pair_->syncWait(lock, pred, pair_->timeout(), "send");
  } else {
    pair_->spinWait(lock, pred);
    auto timeout = pair_->timeout();
    if (timeout == kNoTimeout) {
      sendCv_.wait(lock, pred);
    } else if (!sendCv_.wait_for(lock, timeout, pred)) {
      lock.unlock();
      auto msg = strcat_all("Timed out waiting ", timeout.count(), "ms for send operation to complete");
      pair_->signalExceptionExternal(msg);
      GLB_THROW_TIMEOUT(msg);
    }
  }
  throwIfException();
  if (sendCompletions_ > 0) sendCompletions_--;
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
