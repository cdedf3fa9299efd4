// Opaque, serialisable endpoint identity of a Pair.
// Parity: gloo/transport/address.h:17-27.
#pragma once

#include <string>
#include <vector>

namespace glb {
namespace transport {

class Address {
 public:
  static constexpr size_t kMaxByteSize = 192;
  virtual ~Address() = default;
  virtual std::string str() const = 0;
  virtual std::vector<char> bytes() const = 0;
};

}  // namespace transport
}  // namespace glb
