#include "glb/transport/tcp/address.h"

#include <cstring>

#include "glb/common/logging.h"
#include "glb/transport/tcp/socket.h"

namespace glb {
namespace transport {
namespace tcp {

Address::Address(const struct sockaddr_storage& ss, sequence_number_t seq) {
  std::memset(&impl_, 0, sizeof(impl_));
  impl_.ss = ss;
  impl_.seq = seq;
}

Address::Address(const std::vector<char>& bytes) {
  GLB_ENFORCE_EQ(bytes.size(), sizeof(impl_), "malformed tcp address");
  std::memcpy(&impl_, bytes.data(), sizeof(impl_));
}

std::string Address::str() const { return strcat_all(sockaddrToString(impl_.ss), "#", impl_.seq); }

std::vector<char> Address::bytes() const {
  std::vector<char> out(sizeof(impl_));
  std::memcpy(out.data(), &impl_, sizeof(impl_));
  return out;
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
