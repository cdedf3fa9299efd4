// Bidirectional channel between this process and exactly one peer.
// Parity: gloo/transport/pair.h:21-101.
#pragma once

#include <memory>

#include "glb/transport/address.h"
#include "glb/transport/buffer.h"
#include "glb/transport/unbound_buffer.h"

namespace glb {
namespace transport {

class Pair {
 public:
  virtual ~Pair() = default;

  virtual const Address& address() const = 0;
  virtual void connect(const std::vector<char>& bytes) = 0;
  virtual void close() = 0;
  virtual bool isConnected() = 0;

  // Sync mode: completions are driven by the calling thread instead of the
  // device's I/O thread (lower latency; optionally busy-polling).
  virtual void setSync(bool sync, bool busyPoll) = 0;

  virtual std::unique_ptr<Buffer> createSendBuffer(int slot, void* ptr, size_t size) = 0;
  virtual std::unique_ptr<Buffer> createRecvBuffer(int slot, void* ptr, size_t size) = 0;

  // Unbound operations on this specific pair.
  virtual void send(UnboundBuffer* buf, uint64_t tag, size_t offset, size_t nbytes) = 0;
  virtual void recv(UnboundBuffer* buf, uint64_t tag, size_t offset, size_t nbytes) = 0;

  // Rank of this process among the processes on the same host.
  int getLocalRank() const { return localRank_; }
  void setLocalRank(int r) { localRank_ = r; }

 protected:
  int localRank_ = 0;
};

}  // namespace transport
}  // namespace glb
