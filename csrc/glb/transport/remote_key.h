// Handle to a remotely accessible memory region (one-sided put/get target).
// A key names the owning rank, the region size and a transport-specific token;
// it serialises to text so it can travel through any collective or store.
// Parity: gloo/transport/remote_key.h:8-18 (+ ibverbs/remote_key.{h,cc}).
#pragma once

#include <cstddef>
#include <string>

namespace glb {
namespace transport {

class RemoteKey {
 public:
  RemoteKey(int rank, size_t size) : rank(rank), size(size) {}
  virtual ~RemoteKey() = default;
  virtual std::string serialize() const = 0;

  const int rank;
  const size_t size;
};

}  // namespace transport
}  // namespace glb
