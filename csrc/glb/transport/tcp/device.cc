#include "glb/transport/tcp/device.h"

#include <arpa/inet.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netdb.h>
#include <netinet/in.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "glb/common/linux.h"
#include "glb/common/logging.h"
#include "glb/common/utils.h"
#include "glb/transport/tcp/context.h"

namespace glb {
namespace transport {
namespace tcp {

namespace {

// Fill attr.ai_addr from the first address of `iface` matching the family.
void lookupAddrForIface(struct attr& attr) {
  struct ifaddrs* ifap = nullptr;
  GLB_ENFORCE_NE(getifaddrs(&ifap), -1, "getifaddrs: ", std::strerror(errno));
  bool found = false;
  for (auto* ifa = ifap; ifa != nullptr; ifa = ifa->ifa_next) {
    if (ifa->ifa_addr == nullptr || attr.iface != ifa->ifa_name) continue;
    int fam = ifa->ifa_addr->sa_family;
    if (fam != AF_INET && fam != AF_INET6) continue;
    if (attr.ai_family != AF_UNSPEC && attr.ai_family != fam) continue;
    if (fam == AF_INET6) {
      auto* sa = reinterpret_cast<struct sockaddr_in6*>(ifa->ifa_addr);
      if (IN6_IS_ADDR_LINKLOCAL(&sa->sin6_addr)) continue;  // needs a scope id; skip
    }
    attr.ai_family = fam;
    attr.ai_addrlen = fam == AF_INET ? sizeof(struct sockaddr_in) : sizeof(struct sockaddr_in6);
    std::memcpy(&attr.ai_addr, ifa->ifa_addr, attr.ai_addrlen);
    found = true;
    break;
  }
  freeifaddrs(ifap);
  GLB_ENFORCE(found, "Unable to find an address for interface: ", attr.iface);
}

// Resolve hostname and keep the first address we can actually bind.
void lookupAddrForHostname(struct attr& attr) {
  struct addrinfo hints;
  std::memset(&hints, 0, sizeof(hints));
  hints.ai_family = attr.ai_family;
  hints.ai_socktype = SOCK_STREAM;
  struct addrinfo* result = nullptr;
  int rv = getaddrinfo(attr.hostname.c_str(), nullptr, &hints, &result);
  GLB_ENFORCE(rv == 0, "getaddrinfo(", attr.hostname, "): ", gai_strerror(rv));
  bool found = false;
  for (auto* rp = result; rp != nullptr; rp = rp->ai_next) {
    int fd = ::socket(rp->ai_family, rp->ai_socktype, rp->ai_protocol);
    if (fd == -1) continue;
    rv = ::bind(fd, rp->ai_addr, rp->ai_addrlen);
    ::close(fd);
    if (rv == -1) continue;
    attr.ai_family = rp->ai_family;
    attr.ai_socktype = rp->ai_socktype;
    attr.ai_protocol = rp->ai_protocol;
    std::memcpy(&attr.ai_addr, rp->ai_addr, rp->ai_addrlen);
    attr.ai_addrlen = rp->ai_addrlen;
    found = true;
    break;
  }
  freeaddrinfo(result);
  GLB_ENFORCE(found, "Unable to bind any address of host: ", attr.hostname);
}

// Name of the interface that owns the given address ("" if none).
std::string ifaceForAddr(const struct sockaddr_storage& ss) {
  struct ifaddrs* ifap = nullptr;
  if (getifaddrs(&ifap) == -1) return "";
  std::string out;
  for (auto* ifa = ifap; ifa != nullptr; ifa = ifa->ifa_next) {
    if (ifa->ifa_addr == nullptr || ifa->ifa_addr->sa_family != ss.ss_family) continue;
    bool same = false;
    if (ss.ss_family == AF_INET) {
      same = std::memcmp(&reinterpret_cast<struct sockaddr_in*>(ifa->ifa_addr)->sin_addr,
                         &reinterpret_cast<const struct sockaddr_in*>(&ss)->sin_addr,
                         sizeof(struct in_addr)) == 0;
    } else if (ss.ss_family == AF_INET6) {
      same = std::memcmp(&reinterpret_cast<struct sockaddr_in6*>(ifa->ifa_addr)->sin6_addr,
                         &reinterpret_cast<const struct sockaddr_in6*>(&ss)->sin6_addr,
                         sizeof(struct in6_addr)) == 0;
    }
    if (same) {
      out = ifa->ifa_name;
      break;
    }
  }
  freeifaddrs(ifap);
  return out;
}

struct attr resolve(const struct attr& in) {
  struct attr a = in;
  if (!a.iface.empty()) {
    lookupAddrForIface(a);
  } else {
    if (a.hostname.empty()) a.hostname = getHostname();
    try {
      lookupAddrForHostname(a);
    } catch (const EnforceNotMet&) {
      if (!in.hostname.empty()) throw;
      // The container hostname does not always resolve; loopback always works.
      a.hostname = "127.0.0.1";
      lookupAddrForHostname(a);
    }
  }
  return a;
}

}  // namespace

struct attr resolveAttr(const struct attr& in) { return resolve(in); }

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr& src) {
  return std::make_shared<Device>(resolve(src), /*lazy=*/false);
}

std::shared_ptr<::glb::transport::Device> CreateLazyDevice(const struct attr& src) {
  return std::make_shared<Device>(resolve(src), /*lazy=*/true);
}

// Reads the hello from a freshly accepted socket, then hands it to the device.
class Device::HelloReader : public Handler {
 public:
  HelloReader(Device* dev, Loop* loop, int fd) : dev_(dev), loop_(loop), fd_(fd) {}
  void start() { loop_->registerDescriptor(fd_, EPOLLIN, this); }
  void handleEvents(int events) override {
    while (nread_ < sizeof(Hello)) {
      ssize_t n = ::read(fd_, reinterpret_cast<char*>(&hello_) + nread_, sizeof(Hello) - nread_);
      if (n > 0) {
        nread_ += static_cast<size_t>(n);
        continue;
      }
      if (n == -1 && (errno == EAGAIN || errno == EWOULDBLOCK)) return;
      if (n == -1 && errno == EINTR) continue;
      // EOF or error: drop the connection.
      loop_->unregisterDescriptor(fd_, this);
      ::close(fd_);
      dev_->finishHello(fd_, Hello{0, 0, 0});
      return;
    }
    loop_->unregisterDescriptor(fd_, this);
    dev_->finishHello(fd_, hello_);
  }

 private:
  Device* dev_;
  Loop* loop_;
  int fd_;
  Hello hello_{};
  size_t nread_ = 0;
};

Device::Device(const struct attr& attr, bool lazy) : attr_(attr), lazy_(lazy) {
  int nloops = attr.numLoops > 0 ? attr.numLoops : 1;
  long envLoops = envInt("TCP_LOOPS", 0);
  if (envLoops > 0) nloops = static_cast<int>(envLoops);
  for (int i = 0; i < nloops; i++) loops_.emplace_back(new Loop(attr.portableLoop ? Loop::Backend::POLL : Loop::Backend::EPOLL));

  listener_ = Socket::createForFamily(attr_.ai_family);
  listener_.setReuseAddr(true);
  listener_.setNonBlocking(true);
  int rv = ::bind(listener_.fd(), reinterpret_cast<const struct sockaddr*>(&attr_.ai_addr), attr_.ai_addrlen);
  GLB_ENFORCE_NE(rv, -1, "bind: ", std::strerror(errno));
  rv = ::listen(listener_.fd(), 4096);
  GLB_ENFORCE_NE(rv, -1, "listen: ", std::strerror(errno));
  socklen_t len = sizeof(listenAddr_);
  rv = ::getsockname(listener_.fd(), reinterpret_cast<struct sockaddr*>(&listenAddr_), &len);
  GLB_ENFORCE_NE(rv, -1, "getsockname: ", std::strerror(errno));

  interfaceName_ = attr_.iface.empty() ? ifaceForAddr(listenAddr_) : attr_.iface;
  if (!interfaceName_.empty()) {
    interfaceSpeed_ = getInterfaceSpeedByName(interfaceName_);
    pciBusID_ = interfaceToBusID(interfaceName_);
  }
  loops_[0]->registerDescriptor(listener_.fd(), EPOLLIN, this);
}

Device::~Device() {
  loops_[0]->unregisterDescriptor(listener_.fd(), this);
  {
    std::lock_guard<std::mutex> g(mu_);
    for (auto& kv : readers_) {
      loops_[0]->unregisterDescriptor(kv.first, kv.second.get());
      ::close(kv.first);
    }
    readers_.clear();
    parked_.clear();
    expected_.clear();
  }
  loops_.clear();  // joins the threads
}

std::string Device::str() const {
  return strcat_all("tcp, pci=", pciBusID_, ", iface=", interfaceName_, ", speed=", interfaceSpeed_,
                    ", addr=[", sockaddrToString(listenAddr_), "]");
}

std::shared_ptr<::glb::transport::Context> Device::createContext(int rank, int size) {
  return std::make_shared<Context>(shared_from_this(), rank, size);
}

Address Device::nextAddress() { return Address(listenAddr_, seq_.fetch_add(1)); }
Address Device::addressForSeq(sequence_number_t seq) { return Address(listenAddr_, seq); }

void Device::handleEvents(int /*events*/) {
  while (true) {
    struct sockaddr_storage ss;
    socklen_t len = sizeof(ss);
    int fd = ::accept4(listener_.fd(), reinterpret_cast<struct sockaddr*>(&ss), &len, SOCK_NONBLOCK | SOCK_CLOEXEC);
    if (fd == -1) {
      if (errno == EAGAIN || errno == EWOULDBLOCK) return;
      if (errno == EINTR || errno == ECONNABORTED) continue;
      GLB_WARN("accept: ", std::strerror(errno));
      return;
    }
    auto reader = std::make_unique<HelloReader>(this, loops_[0].get(), fd);
    HelloReader* raw = reader.get();
    {
      std::lock_guard<std::mutex> g(mu_);
      readers_[fd] = std::move(reader);
    }
    raw->start();
  }
}

void Device::finishHello(int fd, const Hello& hello) {
  connect_callback_t cb;
  std::unique_ptr<HelloReader> reader;  // destroyed after we leave its handleEvents frame via defer
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = readers_.find(fd);
    if (it != readers_.end()) {
      reader = std::move(it->second);
      readers_.erase(it);
    }
  }
  // The reader object is still on the call stack (we are inside its handleEvents);
  // free it from the loop once this callback has unwound.
  if (reader) {
    auto* raw = reader.release();
    loops_[0]->defer([raw] { delete raw; });
  }
  if (hello.magic != Hello::kMagic) {
    if (hello.magic != 0) {
      GLB_WARN("dropping connection with bad hello magic");
      ::close(fd);
    }
    return;
  }
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = expected_.find(hello.seq);
    if (it == expected_.end()) {
      parked_.emplace(hello.seq, Socket(fd));
      return;
    }
    cb = std::move(it->second);
    expected_.erase(it);
  }
  cb(Socket(fd));
}

void Device::expectConnection(sequence_number_t seq, connect_callback_t cb) {
  Socket sock;
  {
    std::lock_guard<std::mutex> g(mu_);
    auto it = parked_.find(seq);
    if (it == parked_.end()) {
      expected_[seq] = std::move(cb);
      return;
    }
    sock = std::move(it->second);
    parked_.erase(it);
  }
  cb(std::move(sock));
}

void Device::cancelExpectation(sequence_number_t seq) {
  std::lock_guard<std::mutex> g(mu_);
  expected_.erase(seq);
  parked_.erase(seq);
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
