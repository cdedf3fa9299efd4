// Unbound buffer: a (ptr, size) region that is not tied to a Pair. It can be the
// source of sends to any rank, the destination of receives from a given rank or
// from "any of these ranks", and (when the transport supports it) the target or
// origin of one-sided put / get through a RemoteKey.
//
// Rules (same as the reference, unbound_buffer.h:24-36): at most one pending
// operation of each kind is assumed by the collectives; waitSend / waitRecv return
// false when aborted; a timeout poisons the whole context and throws IoException.
// Parity: gloo/transport/unbound_buffer.h:36-153.
#pragma once

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <vector>

#include "glb/common/error.h"
#include "glb/transport/remote_key.h"

namespace glb {
namespace transport {

class UnboundBuffer {
 public:
  UnboundBuffer(void* ptr, size_t size) : ptr(ptr), size(size) {}
  virtual ~UnboundBuffer() = default;

  void* const ptr;
  const size_t size;

  static constexpr size_t kUnspecifiedByteCount = std::numeric_limits<size_t>::max();

  // If `rank` is given it receives the peer of the completed operation.
  // Returns true on completion, false when aborted.
  virtual bool waitRecv(int* rank, std::chrono::milliseconds timeout) = 0;
  virtual bool waitSend(int* rank, std::chrono::milliseconds timeout) = 0;
  virtual void abortWaitRecv() = 0;
  virtual void abortWaitSend() = 0;

  bool waitRecv() { return waitRecv(nullptr, kUnsetTimeout); }
  bool waitSend() { return waitSend(nullptr, kUnsetTimeout); }
  bool waitRecv(int* rank) { return waitRecv(rank, kUnsetTimeout); }
  bool waitSend(int* rank) { return waitSend(rank, kUnsetTimeout); }
  bool waitRecv(std::chrono::milliseconds timeout) { return waitRecv(nullptr, timeout); }
  bool waitSend(std::chrono::milliseconds timeout) { return waitSend(nullptr, timeout); }

  virtual void send(int dstRank, uint64_t slot, size_t offset = 0,
                    size_t nbytes = kUnspecifiedByteCount) = 0;
  virtual void recv(int srcRank, uint64_t slot, size_t offset = 0,
                    size_t nbytes = kUnspecifiedByteCount) = 0;
  // Receive from whichever of `srcRanks` sends first on `slot`.
  virtual void recv(std::vector<int> srcRanks, uint64_t slot, size_t offset = 0,
                    size_t nbytes = kUnspecifiedByteCount) = 0;

  // One-sided access. getRemoteKey() exposes this buffer to peers; put() writes
  // local bytes [offset, offset+nbytes) to the remote region at roffset (complete
  // after waitSend); get() reads remote bytes into the local region (complete
  // after waitRecv). Both are bounds-checked against the key.
  virtual std::unique_ptr<RemoteKey> getRemoteKey() const {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("getRemoteKey() not supported by this transport");
  }
  virtual void put(const RemoteKey& key, uint64_t slot, size_t offset, size_t roffset, size_t nbytes) {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("put(RemoteKey) not supported by this transport");
  }
  virtual void get(const RemoteKey& key, uint64_t slot, size_t offset, size_t roffset, size_t nbytes) {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("get(RemoteKey) not supported by this transport");
  }
};

}  // namespace transport
}  // namespace glb
