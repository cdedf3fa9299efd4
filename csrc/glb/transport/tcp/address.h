// tcp::Address = socket address of the owning Device's listener + the sequence
// number that identifies one Pair behind that listener.
// Parity: gloo/transport/tcp/address.{h,cc}.
#pragma once

#include <sys/socket.h>

#include <cstdint>

#include "glb/transport/address.h"

namespace glb {
namespace transport {
namespace tcp {

using sequence_number_t = uint64_t;

class Address : public ::glb::transport::Address {
 public:
  Address() = default;
  Address(const struct sockaddr_storage& ss, sequence_number_t seq);
  explicit Address(const std::vector<char>& bytes);

  std::string str() const override;
  std::vector<char> bytes() const override;

  const struct sockaddr_storage& sockaddr() const { return impl_.ss; }
  sequence_number_t seq() const { return impl_.seq; }

 private:
  struct Impl {
    struct sockaddr_storage ss;
    sequence_number_t seq;
  };
  static_assert(sizeof(Impl) <= kMaxByteSize, "address too large");
  Impl impl_{};
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
