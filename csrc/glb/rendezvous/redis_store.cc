#include "glb/rendezvous/redis_store.h"

#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <thread>

#include "glb/common/logging.h"

namespace glb {
namespace rendezvous {

RedisStore::RedisStore(const std::string& host, int port) : host_(host), port_(port) {
  struct addrinfo hints;
  std::memset(&hints, 0, sizeof(hints));
  hints.ai_family = AF_UNSPEC;
  hints.ai_socktype = SOCK_STREAM;
  struct addrinfo* res = nullptr;
  int rv = getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res);
  if (rv != 0) GLB_THROW_IO_EXCEPTION("Connecting to Redis: getaddrinfo(", host, "): ", gai_strerror(rv));
  std::string err = "no address";
  for (auto* rp = res; rp != nullptr; rp = rp->ai_next) {
    int fd = ::socket(rp->ai_family, rp->ai_socktype | SOCK_CLOEXEC, rp->ai_protocol);
    if (fd < 0) continue;
    if (::connect(fd, rp->ai_addr, rp->ai_addrlen) == 0) {
      fd_ = fd;
      break;
    }
    err = std::strerror(errno);
    ::close(fd);
  }
  freeaddrinfo(res);
  if (fd_ < 0) GLB_THROW_IO_EXCEPTION("Connecting to Redis (", host, ":", port, "): ", err);
  int one = 1;
  ::setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
}

RedisStore::~RedisStore() {
  if (fd_ >= 0) ::close(fd_);
}

void RedisStore::readExact(char* dst, size_t n) {
  while (rbuf_.size() < n) {
    char tmp[4096];
    ssize_t r = ::read(fd_, tmp, sizeof(tmp));
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) GLB_THROW_IO_EXCEPTION("Redis connection lost");
    rbuf_.append(tmp, static_cast<size_t>(r));
  }
  std::memcpy(dst, rbuf_.data(), n);
  rbuf_.erase(0, n);
}

std::string RedisStore::readLine() {
  while (true) {
    auto pos = rbuf_.find("\r\n");
    if (pos != std::string::npos) {
      std::string line = rbuf_.substr(0, pos);
      rbuf_.erase(0, pos + 2);
      return line;
    }
    char tmp[4096];
    ssize_t r = ::read(fd_, tmp, sizeof(tmp));
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) GLB_THROW_IO_EXCEPTION("Redis connection lost");
    rbuf_.append(tmp, static_cast<size_t>(r));
  }
}

RedisStore::Reply RedisStore::readReply() {
  Reply rep;
  std::string line = readLine();
  GLB_ENFORCE(!line.empty(), "empty Redis reply");
  rep.type = line[0];
  std::string rest = line.substr(1);
  switch (rep.type) {
    case '+': rep.str = rest; break;
    case '-': GLB_THROW_IO_EXCEPTION("Redis error: ", rest);
    case ':': rep.integer = std::stoll(rest); break;
    case '$': {
      long len = std::stol(rest);
      if (len < 0) {
        rep.nil = true;
      } else {
        rep.str.resize(static_cast<size_t>(len));
        if (len > 0) readExact(&rep.str[0], static_cast<size_t>(len));
        char crlf[2];
        readExact(crlf, 2);
      }
      break;
    }
    case '*': {
      long n = std::stol(rest);
      for (long i = 0; i < n; i++) rep.elems.push_back(readReply());
      break;
    }
    default: GLB_THROW_IO_EXCEPTION("unexpected Redis reply type '", rep.type, "'");
  }
  return rep;
}

RedisStore::Reply RedisStore::command(const std::vector<std::string>& args) {
  std::string req = "*" + std::to_string(args.size()) + "\r\n";
  for (const auto& a : args) {
    req += "$" + std::to_string(a.size()) + "\r\n";
    req += a;
    req += "\r\n";
  }
  std::lock_guard<std::mutex> g(mu_);
  size_t off = 0;
  while (off < req.size()) {
    ssize_t n = ::send(fd_, req.data() + off, req.size() - off, MSG_NOSIGNAL);
    if (n < 0 && errno == EINTR) continue;
    if (n <= 0) GLB_THROW_IO_EXCEPTION("Redis write failed: ", std::strerror(errno));
    off += static_cast<size_t>(n);
  }
  return readReply();
}

void RedisStore::set(const std::string& key, const Bytes& data) {
  auto rep = command({"SETNX", key, std::string(data.begin(), data.end())});
  GLB_ENFORCE_EQ(rep.integer, 1, "Key '", key, "' already set");
}

IStore::Bytes RedisStore::get(const std::string& key) {
  wait({key}, kDefaultTimeout);
  auto rep = command({"GET", key});
  GLB_ENFORCE(!rep.nil, "Key '", key, "' not set");
  return Bytes(rep.str.begin(), rep.str.end());
}

bool RedisStore::check(const std::vector<std::string>& keys) {
  std::vector<std::string> args{"EXISTS"};
  args.insert(args.end(), keys.begin(), keys.end());
  auto rep = command(args);
  return rep.integer == static_cast<int64_t>(keys.size());
}

void RedisStore::wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) {
  const auto start = std::chrono::steady_clock::now();
  while (!check(keys)) {
    if (timeout != kNoTimeout && std::chrono::steady_clock::now() - start > timeout) {
      GLB_THROW_IO_EXCEPTION("Wait timeout for key(s): ", keys.size() == 1 ? keys[0] : "[multiple]");
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(10));
  }
}

std::vector<IStore::Bytes> RedisStore::multi_get(const std::vector<std::string>& keys) {
  wait(keys, kDefaultTimeout);
  std::vector<std::string> args{"MGET"};
  args.insert(args.end(), keys.begin(), keys.end());
  auto rep = command(args);
  std::vector<Bytes> out;
  for (auto& e : rep.elems) out.emplace_back(e.str.begin(), e.str.end());
  return out;
}

void RedisStore::multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
  GLB_ENFORCE_EQ(keys.size(), values.size());
  for (size_t i = 0; i < keys.size(); i++) set(keys[i], values[i]);
}

void RedisStore::append(const std::string& key, const Bytes& data) {
  command({"APPEND", key, std::string(data.begin(), data.end())});
}

int64_t RedisStore::add(const std::string& key, int64_t value) {
  return command({"INCRBY", key, std::to_string(value)}).integer;
}

}  // namespace rendezvous
}  // namespace glb
