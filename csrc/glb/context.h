// glb::Context — a communicator: (rank, size, base) plus the transport context
// holding one Pair per peer. No global or thread-local state: several contexts
// can coexist in a process (that is how threads-as-ranks tests work).
// Parity: gloo/context.{h,cc}.
#pragma once

#include <atomic>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "glb/transport/pair.h"

namespace glb {

namespace transport {
class Context;
class Device;
class UnboundBuffer;
}  // namespace transport

class Context : public std::enable_shared_from_this<Context> {
 public:
  Context(int rank, int size, int base = 2);
  virtual ~Context();

  const int rank;
  const int size;
  int base;  // fan-out for the bcube algorithms

  std::shared_ptr<transport::Device>& getDevice();
  std::unique_ptr<transport::Pair>& getPair(int i);
  std::shared_ptr<transport::Context>& getTransportContext() { return transportContext_; }

  // Factory for unbound buffers (see transport/unbound_buffer.h).
  std::unique_ptr<transport::UnboundBuffer> createUnboundBuffer(void* ptr, size_t size);

  // Reserve `numSlots` consecutive slot ids for an old-style algorithm instance.
  int nextSlot(int numSlots = 1);

  void closeConnections();

  void setTimeout(std::chrono::milliseconds timeout);
  std::chrono::milliseconds getTimeout() const;

  std::unique_ptr<transport::RemoteKey> deserializeRemoteKey(const std::string& serialized);

  // Objects whose lifetime is tied to this context (e.g. the CUDA PeerContext that the
  // old-style CUDA algorithms share). They are destroyed before the transport.
  std::shared_ptr<void> getAttachment(const std::string& key);
  void setAttachment(const std::string& key, std::shared_ptr<void> value);
  void clearAttachments();

 protected:
  std::shared_ptr<transport::Device> device_;
  std::shared_ptr<transport::Context> transportContext_;
  std::atomic<int> slot_{0};
  std::chrono::milliseconds timeout_;
  std::mutex attachMu_;
  std::map<std::string, std::shared_ptr<void>> attachments_;  // declared last: destroyed first
};

}  // namespace glb
