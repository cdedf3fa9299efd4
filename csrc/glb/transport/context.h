// Per-communicator transport state: the vector of Pairs (one per peer), the
// rendezvous that connects them through a Store, and the UnboundBuffer factory.
// Parity: gloo/transport/context.h:36-298 — the recv-from-any bookkeeping that the
// reference keeps here (Tally/Mutator) lives in tcp::Context, because with our
// eager wire protocol matching is purely receiver-local.
#pragma once

#include <chrono>
#include <memory>
#include <mutex>
#include <vector>

#include "glb/common/store.h"
#include "glb/transport/pair.h"
#include "glb/transport/remote_key.h"
#include "glb/transport/unbound_buffer.h"

namespace glb {
class Context;  // the communicator that owns a transport context
namespace transport {

class Context {
 public:
  Context(int rank, int size) : rank(rank), size(size), pairs_(size) {}
  virtual ~Context() = default;

  const int rank;
  const int size;

  virtual std::unique_ptr<Pair>& getPair(int rank) { return pairs_.at(rank); }
  virtual std::unique_ptr<Pair>& createPair(int rank) = 0;
  // The pair object as is (no lazy connect); may be null.
  virtual Pair* peekPair(int rank) { return pairs_.at(rank).get(); }
  // Called once the owning glb::Context holds this transport context (end of
  // connectFullMesh / ContextFactory::makeContext, i.e. collectively on every rank).
  virtual void onAttached(const std::weak_ptr<::glb::Context>& /*owner*/) {}

  // Generic rendezvous: every rank publishes the addresses of all its pairs under
  // its rank key, then connects to each peer (O(P^2) store traffic). Transports
  // may override with something leaner (tcp does).
  virtual void createAndConnectAllPairs(std::shared_ptr<IStore> store);

  virtual std::unique_ptr<UnboundBuffer> createUnboundBuffer(void* ptr, size_t size) = 0;

  virtual std::unique_ptr<RemoteKey> deserializeRemoteKey(const std::string& serialized) {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("this transport has no one-sided support");
  }

  virtual void setTimeout(std::chrono::milliseconds timeout) { timeout_ = timeout; }
  std::chrono::milliseconds getTimeout() const { return timeout_; }

  // Address blob used by ContextFactory to wire up a derived context without a store.
  virtual std::vector<char> exportRendezvousBlob() {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("exportRendezvousBlob unsupported");
  }
  virtual void connectWithBlobs(const std::vector<std::vector<char>>& blobs) {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("connectWithBlobs unsupported");
  }

 protected:
  std::vector<std::unique_ptr<Pair>> pairs_;
  std::chrono::milliseconds timeout_{std::chrono::seconds(30)};
  std::mutex mutex_;
};

}  // namespace transport
}  // namespace glb
