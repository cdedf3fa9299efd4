#include "glb/context.h"

#include "glb/common/logging.h"
#include "glb/transport/context.h"
#include "glb/transport/device.h"
#include "glb/transport/unbound_buffer.h"

namespace glb {

static constexpr std::chrono::milliseconds kTimeoutDefault = std::chrono::seconds(30);

Context::Context(int rank, int size, int base) : rank(rank), size(size), base(base), timeout_(kTimeoutDefault) {
  GLB_ENFORCE_GE(rank, 0);
  GLB_ENFORCE_LT(rank, size);
  GLB_ENFORCE_GE(size, 1);
}

Context::~Context() { clearAttachments(); }

std::shared_ptr<void> Context::getAttachment(const std::string& key) {
  std::lock_guard<std::mutex> g(attachMu_);
  auto it = attachments_.find(key);
  return it == attachments_.end() ? nullptr : it->second;
}

void Context::setAttachment(const std::string& key, std::shared_ptr<void> value) {
  std::lock_guard<std::mutex> g(attachMu_);
  attachments_[key] = std::move(value);
}

void Context::clearAttachments() {
  std::map<std::string, std::shared_ptr<void>> drop;
  {
    std::lock_guard<std::mutex> g(attachMu_);
    drop.swap(attachments_);
  }
}

std::shared_ptr<transport::Device>& Context::getDevice() {
  GLB_ENFORCE(device_, "Device not set!");
  return device_;
}

std::unique_ptr<transport::Pair>& Context::getPair(int i) {
  GLB_ENFORCE(transportContext_, "Transport context not set!");
  return transportContext_->getPair(i);
}

std::unique_ptr<transport::UnboundBuffer> Context::createUnboundBuffer(void* ptr, size_t size) {
  GLB_ENFORCE(transportContext_, "Transport context not set!");
  return transportContext_->createUnboundBuffer(ptr, size);
}

int Context::nextSlot(int numSlots) {
  GLB_ENFORCE_GT(numSlots, 0);
  return slot_.fetch_add(numSlots);
}

void Context::closeConnections() {
  clearAttachments();
  if (!transportContext_) return;
  for (int i = 0; i < size; i++) {
    auto* pair = transportContext_->peekPair(i);
    if (pair != nullptr) pair->close();
  }
}

void Context::setTimeout(std::chrono::milliseconds timeout) {
  GLB_ENFORCE(timeout.count() >= 0, "Invalid timeout: ", timeout.count());
  timeout_ = timeout;
  if (transportContext_) transportContext_->setTimeout(timeout);
}

std::chrono::milliseconds Context::getTimeout() const { return timeout_; }

std::unique_ptr<transport::RemoteKey> Context::deserializeRemoteKey(const std::string& serialized) {
  GLB_ENFORCE(transportContext_, "Transport context not set!");
  return transportContext_->deserializeRemoteKey(serialized);
}

}  // namespace glb
