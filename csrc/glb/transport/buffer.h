// Bound buffer: (slot, ptr, size) registered with one specific Pair at creation.
// A send buffer's send(offset, length, roffset) is a one-sided write into the
// peer's recv buffer registered under the same slot; completions are counted and
// consumed by waitSend / waitRecv. Used by the old-style Algorithm classes.
// Parity: gloo/transport/buffer.h:16-41.
#pragma once

#include <cstddef>

namespace glb {
namespace transport {

class Buffer {
 public:
  Buffer(int slot, void* ptr, size_t size) : slot_(slot), ptr_(ptr), size_(size), debug_(false) {}
  virtual ~Buffer() = default;

  virtual void setDebug(bool debug) { debug_ = debug; }

  virtual void send(size_t offset, size_t length, size_t roffset = 0) = 0;
  // Send entire buffer by default.
  void send() { send(0, size_); }

  virtual void waitRecv() = 0;
  virtual void waitSend() = 0;

  int slot() const { return slot_; }
  void* ptr() const { return ptr_; }
  size_t size() const { return size_; }

 protected:
  int slot_;
  void* ptr_;
  size_t size_;
  bool debug_;
};

}  // namespace transport
}  // namespace glb
