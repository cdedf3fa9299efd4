// Device creation attributes: pick the listening address by hostname or by
// network interface. Parity: gloo/transport/tcp/attr.h:19-40.
#pragma once

#include <sys/socket.h>

#include <string>

namespace glb {
namespace transport {
namespace tcp {

struct attr {
  attr() = default;
  /* implicit */ attr(const char* host) : hostname(host) {}
  /* implicit */ attr(const std::string& host) : hostname(host) {}

  std::string hostname;  // name or literal address to bind; empty -> this host's name
  std::string iface;     // interface name; wins over hostname when set
  int ai_family = AF_UNSPEC;

  // Resolved by CreateDevice.
  int ai_socktype = SOCK_STREAM;
  int ai_protocol = 0;
  struct sockaddr_storage ai_addr {};
  socklen_t ai_addrlen = 0;

  // Number of event-loop threads; pairs are sharded across them round-robin.
  int numLoops = 1;
  // Readiness backend of the loops: epoll (Linux) or the portable poll(2) reactor that
  // the "uv" device uses (the role libuv plays in the reference: any POSIX system).
  bool portableLoop = false;
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
