#include "glb/transport/tcp/socket.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "glb/common/logging.h"

namespace glb {
namespace transport {
namespace tcp {

Socket& Socket::operator=(Socket&& o) noexcept {
  if (this != &o) {
    close();
    fd_ = o.fd_;
    o.fd_ = -1;
  }
  return *this;
}

Socket::~Socket() { close(); }

Socket Socket::createForFamily(int family) {
  int fd = ::socket(family, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (fd < 0) GLB_THROW_IO_EXCEPTION("socket: ", std::strerror(errno));
  return Socket(fd);
}

int Socket::release() {
  int fd = fd_;
  fd_ = -1;
  return fd;
}

void Socket::close() {
  if (fd_ >= 0) {
    ::close(fd_);
    fd_ = -1;
  }
}

void Socket::setNonBlocking(bool on) {
  int flags = ::fcntl(fd_, F_GETFL);
  GLB_ENFORCE_NE(flags, -1, "fcntl: ", std::strerror(errno));
  flags = on ? (flags | O_NONBLOCK) : (flags & ~O_NONBLOCK);
  GLB_ENFORCE_NE(::fcntl(fd_, F_SETFL, flags), -1, "fcntl: ", std::strerror(errno));
}

void Socket::setNoDelay(bool on) {
  int v = on ? 1 : 0;
  ::setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &v, sizeof(v));
}

void Socket::setReuseAddr(bool on) {
  int v = on ? 1 : 0;
  ::setsockopt(fd_, SOL_SOCKET, SO_REUSEADDR, &v, sizeof(v));
}

void Socket::setLingerZero() {
  struct linger sl;
  sl.l_onoff = 1;
  sl.l_linger = 0;
  ::setsockopt(fd_, SOL_SOCKET, SO_LINGER, &sl, sizeof(sl));
}

void Socket::setTimeouts(std::chrono::milliseconds t) {
  struct timeval tv;
  tv.tv_sec = t.count() / 1000;
  tv.tv_usec = (t.count() % 1000) * 1000;
  ::setsockopt(fd_, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  ::setsockopt(fd_, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
}

void Socket::growBuffers(int bytes) {
  ::setsockopt(fd_, SOL_SOCKET, SO_SNDBUF, &bytes, sizeof(bytes));
  ::setsockopt(fd_, SOL_SOCKET, SO_RCVBUF, &bytes, sizeof(bytes));
}

std::string sockaddrToString(const struct sockaddr_storage& ss) {
  char host[INET6_ADDRSTRLEN] = {0};
  int port = 0;
  if (ss.ss_family == AF_INET) {
    auto* in = reinterpret_cast<const struct sockaddr_in*>(&ss);
    inet_ntop(AF_INET, &in->sin_addr, host, sizeof(host));
    port = ntohs(in->sin_port);
    return std::string(host) + ":" + std::to_string(port);
  }
  if (ss.ss_family == AF_INET6) {
    auto* in6 = reinterpret_cast<const struct sockaddr_in6*>(&ss);
    inet_ntop(AF_INET6, &in6->sin6_addr, host, sizeof(host));
    port = ntohs(in6->sin6_port);
    return "[" + std::string(host) + "]:" + std::to_string(port);
  }
  return "<unknown family>";
}

socklen_t sockaddrLen(const struct sockaddr_storage& ss) {
  return ss.ss_family == AF_INET6 ? sizeof(struct sockaddr_in6) : sizeof(struct sockaddr_in);
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
