// TLS-protected TCP transport: mutual authentication (both sides present a
// certificate signed by the given CA), TLS >= 1.2, security level 2 — the reference's
// posture (tls/context.cc:51-134). Device / Context / Pair subclass their tcp
// counterparts; only the byte-level I/O and the post-connect handshake differ.
// Parity: gloo/transport/tcp/tls/{device,context,pair}.{h,cc}.
#pragma once

#include <memory>
#include <string>

#include "glb/transport/tcp/context.h"
#include "glb/transport/tcp/device.h"
#include "glb/transport/tcp/pair.h"
#include "glb/transport/tcp/tls/openssl.h"

namespace glb {
namespace transport {
namespace tcp {
namespace tls {

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr& src, const std::string& pkey_file,
                                                      const std::string& cert_file, const std::string& ca_file,
                                                      const std::string& ca_path);

class Device : public ::glb::transport::tcp::Device {
 public:
  Device(const struct attr& attr, std::string pkey, std::string cert, std::string caFile, std::string caPath);
  ~Device() override;
  std::string str() const override;
  std::shared_ptr<::glb::transport::Context> createContext(int rank, int size) override;
  SSL_CTX* sslCtx() const { return ctx_; }

 private:
  SSL_CTX* ctx_ = nullptr;
};

class Context : public ::glb::transport::tcp::Context {
 public:
  Context(std::shared_ptr<Device> device, int rank, int size);
  std::unique_ptr<::glb::transport::Pair>& createPair(int rank) override;

 private:
  std::shared_ptr<Device> tlsDevice_;
};

class Pair : public ::glb::transport::tcp::Pair {
 public:
  Pair(::glb::transport::tcp::Context* context, Device* device, int selfRank, int peerRank,
       std::chrono::milliseconds timeout, bool lazy);
  ~Pair() override;

 protected:
  ssize_t ioRecv(void* buf, size_t len) override;
  ssize_t ioSend(const struct iovec* iov, int iovcnt) override;
  void ioHandshake(bool isInitiator) override;
  void ioShutdown() override;
  bool allowCma() const override { return false; }  // payloads stay inside the TLS session
  bool ioPending() override { return ssl_ != nullptr && openssl().SSL_pending(ssl_) > 0; }

 private:
  Device* tlsDevice_;
  SSL* ssl_ = nullptr;
};

}  // namespace tls
}  // namespace tcp
}  // namespace transport
}  // namespace glb
