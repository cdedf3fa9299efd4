// Thin RAII + option helpers over BSD sockets.
// Parity: gloo/transport/tcp/socket.{h,cc}.
#pragma once

#include <sys/socket.h>

#include <chrono>
#include <string>

namespace glb {
namespace transport {
namespace tcp {

class Socket {
 public:
  Socket() = default;
  explicit Socket(int fd) : fd_(fd) {}
  Socket(Socket&& o) noexcept : fd_(o.fd_) { o.fd_ = -1; }
  Socket& operator=(Socket&& o) noexcept;
  Socket(const Socket&) = delete;
  Socket& operator=(const Socket&) = delete;
  ~Socket();

  static Socket createForFamily(int family);

  int fd() const { return fd_; }
  bool valid() const { return fd_ >= 0; }
  int release();
  void close();

  void setNonBlocking(bool on);
  void setNoDelay(bool on);
  void setReuseAddr(bool on);
  void setLingerZero();
  void setTimeouts(std::chrono::milliseconds t);  // SO_RCVTIMEO / SO_SNDTIMEO (sync mode)
  void growBuffers(int bytes);                    // best-effort SO_SNDBUF / SO_RCVBUF

 private:
  int fd_ = -1;
};

std::string sockaddrToString(const struct sockaddr_storage& ss);
socklen_t sockaddrLen(const struct sockaddr_storage& ss);

}  // namespace tcp
}  // namespace transport
}  // namespace glb
