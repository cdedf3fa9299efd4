#include "glb/transport/tcp/tls/device.h"

#include <csignal>

#include <poll.h>
#include <sys/uio.h>

#include <cerrno>

#include "glb/common/logging.h"

namespace glb {
namespace transport {
namespace tcp {

// Defined in tcp/device.cc (attribute resolution shared with the plain transport).
struct attr resolveAttr(const struct attr& in);

namespace tls {

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr& src, const std::string& pkey_file,
                                                      const std::string& cert_file, const std::string& ca_file,
                                                      const std::string& ca_path) {
  GLB_ENFORCE(!pkey_file.empty(), "tls: private key file is required");
  GLB_ENFORCE(!cert_file.empty(), "tls: certificate file is required");
  GLB_ENFORCE(!ca_file.empty() || !ca_path.empty(), "tls: a CA file or CA path is required");
  return std::make_shared<Device>(resolveAttr(src), pkey_file, cert_file, ca_file, ca_path);
}

Device::Device(const struct attr& attr, std::string pkey, std::string cert, std::string caFile, std::string caPath)
    : ::glb::transport::tcp::Device(attr, /*lazy=*/false) {
  // OpenSSL writes through plain write(2), which cannot carry MSG_NOSIGNAL the way the
  // tcp transport's sendmsg does: a peer that went away would kill the process with
  // SIGPIPE instead of surfacing as an IoException. Only the default disposition is
  // replaced; an application handler is left alone.
  {
    struct sigaction cur;
    if (::sigaction(SIGPIPE, nullptr, &cur) == 0 && cur.sa_handler == SIG_DFL) ::signal(SIGPIPE, SIG_IGN);
  }
  const auto& o = openssl();
  ctx_ = o.SSL_CTX_new(o.TLS_method());
  GLB_ENFORCE(ctx_ != nullptr, "SSL_CTX_new: ", opensslLastError());
  auto fail = [&](const char* what) {
    std::string err = opensslLastError();
    o.SSL_CTX_free(ctx_);
    ctx_ = nullptr;
    GLB_THROW(::glb::IoException, "tls: ", what, ": ", err);
  };
  o.SSL_CTX_ctrl(ctx_, SSL_CTRL_SET_MIN_PROTO_VERSION, TLS1_2_VERSION, nullptr);
  o.SSL_CTX_set_security_level(ctx_, 2);
  o.SSL_CTX_ctrl(ctx_, SSL_CTRL_MODE, SSL_MODE_ENABLE_PARTIAL_WRITE | SSL_MODE_ACCEPT_MOVING_WRITE_BUFFER, nullptr);
  if (o.SSL_CTX_use_certificate_chain_file(ctx_, cert.c_str()) != 1) fail("loading certificate");
  if (o.SSL_CTX_use_PrivateKey_file(ctx_, pkey.c_str(), SSL_FILETYPE_PEM) != 1) fail("loading private key");
  if (o.SSL_CTX_check_private_key(ctx_) != 1) fail("private key does not match the certificate");
  if (o.SSL_CTX_load_verify_locations(ctx_, caFile.empty() ? nullptr : caFile.c_str(),
                                      caPath.empty() ? nullptr : caPath.c_str()) != 1) {
    fail("loading CA");
  }
  // Mutual authentication: both sides must present a certificate signed by the CA.
  o.SSL_CTX_set_verify(ctx_, SSL_VERIFY_PEER | SSL_VERIFY_FAIL_IF_NO_PEER_CERT, nullptr);
}

Device::~Device() {
  if (ctx_ != nullptr) openssl().SSL_CTX_free(ctx_);
}

std::string Device::str() const { return "tls+" + ::glb::transport::tcp::Device::str(); }

std::shared_ptr<::glb::transport::Context> Device::createContext(int rank, int size) {
  return std::make_shared<Context>(std::static_pointer_cast<Device>(shared_from_this()), rank, size);
}

Context::Context(std::shared_ptr<Device> device, int rank, int size)
    : ::glb::transport::tcp::Context(device, rank, size), tlsDevice_(std::move(device)) {}

std::unique_ptr<::glb::transport::Pair>& Context::createPair(int peer) {
  GLB_ENFORCE(peer >= 0 && peer < size && peer != rank, "invalid peer rank ", peer);
  pairs_[peer] = std::make_unique<Pair>(this, tlsDevice_.get(), rank, peer, getTimeout(), false);
  return pairs_[peer];
}

Pair::Pair(::glb::transport::tcp::Context* context, Device* device, int selfRank, int peerRank,
           std::chrono::milliseconds timeout, bool lazy)
    : ::glb::transport::tcp::Pair(context, device, selfRank, peerRank, timeout, lazy), tlsDevice_(device) {}

Pair::~Pair() {
  // The loop thread may be inside SSL_read on this session right now; the base class only
  // detaches from it in its own destructor, i.e. after this one.
  quiesce();
  if (ssl_ != nullptr) {
    if (fd() >= 0) openssl().SSL_shutdown(ssl_);  // close_notify, best effort (the base closes the fd)
    openssl().SSL_free(ssl_);
    ssl_ = nullptr;
  }
}

// Runs right after the TCP connection is attached, before any framed traffic. The
// socket is non-blocking: drive the handshake with poll().
void Pair::ioHandshake(bool isInitiator) {
  const auto& o = openssl();
  ssl_ = o.SSL_new(tlsDevice_->sslCtx());
  GLB_ENFORCE(ssl_ != nullptr, "SSL_new: ", opensslLastError());
  GLB_ENFORCE_EQ(o.SSL_set_fd(ssl_, fd()), 1, "SSL_set_fd: ", opensslLastError());
  if (isInitiator) {
    o.SSL_set_connect_state(ssl_);
  } else {
    o.SSL_set_accept_state(ssl_);
  }
  const auto start = std::chrono::steady_clock::now();
  const auto limit = timeout() == kNoTimeout ? std::chrono::milliseconds(60000) : timeout();
  while (true) {
    int rv = o.SSL_do_handshake(ssl_);
    if (rv == 1) return;
    int err = o.SSL_get_error(ssl_, rv);
    if (err != SSL_ERROR_WANT_READ && err != SSL_ERROR_WANT_WRITE) {
      GLB_THROW(::glb::IoException, "TLS handshake failed: ", opensslLastError());
    }
    if (std::chrono::steady_clock::now() - start > limit) GLB_THROW(::glb::IoException, "TLS handshake timed out");
    struct pollfd pfd = {fd(), static_cast<short>(err == SSL_ERROR_WANT_READ ? POLLIN : POLLOUT), 0};
    ::poll(&pfd, 1, 50);
  }
}

ssize_t Pair::ioRecv(void* buf, size_t len) {
  const auto& o = openssl();
  if (len == 0) return 0;
  // SSL_get_error consults the calling thread's error queue and errno; reads come from
  // the loop thread and from spinning waiters, so both start clean every time.
  o.ERR_clear_error();
  errno = 0;
  int n = o.SSL_read(ssl_, buf, static_cast<int>(std::min<size_t>(len, 1 << 30)));
  if (n > 0) return n;
  int err = o.SSL_get_error(ssl_, n);
  if (err == SSL_ERROR_WANT_READ || err == SSL_ERROR_WANT_WRITE) {
    errno = EAGAIN;
    return -1;
  }
  if (err == SSL_ERROR_ZERO_RETURN) return 0;
  if (err == SSL_ERROR_SYSCALL && errno == 0) return 0;  // peer vanished without close_notify
  if (errno == 0) errno = EIO;
  return -1;
}

ssize_t Pair::ioSend(const struct iovec* iov, int iovcnt) {
  const auto& o = openssl();
  ssize_t total = 0;
  for (int i = 0; i < iovcnt; i++) {
    size_t off = 0;
    while (off < iov[i].iov_len) {
      o.ERR_clear_error();
      errno = 0;
      int n = o.SSL_write(ssl_, static_cast<const char*>(iov[i].iov_base) + off,
                          static_cast<int>(std::min<size_t>(iov[i].iov_len - off, 1 << 30)));
      if (n > 0) {
        off += static_cast<size_t>(n);
        total += n;
        continue;
      }
      int err = o.SSL_get_error(ssl_, n);
      if (err == SSL_ERROR_WANT_READ || err == SSL_ERROR_WANT_WRITE) {
        if (total > 0) return total;
        errno = EAGAIN;
        return -1;
      }
      if (total > 0) return total;
      if (errno == 0) errno = EIO;
      return -1;
    }
  }
  return total;
}

void Pair::ioShutdown() {
  if (ssl_ != nullptr) openssl().SSL_shutdown(ssl_);
}

}  // namespace tls
}  // namespace tcp
}  // namespace transport
}  // namespace glb
