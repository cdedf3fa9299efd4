// Run-time loaded OpenSSL (libssl.so.3 / libssl.so.1.1): the library has no link-time
// dependency on OpenSSL; tls::CreateDevice throws if it cannot be loaded.
// Parity: gloo/transport/tcp/tls/openssl.{h,cc} + dynamic_library.{h,cc}
// (the reference's USE_TCP_OPENSSL_LOAD mode).
#pragma once

#include <openssl/err.h>
#include <openssl/ssl.h>

#include <string>

namespace glb {
namespace transport {
namespace tcp {
namespace tls {

struct OpenSSL {
  const SSL_METHOD* (*TLS_method)();
  SSL_CTX* (*SSL_CTX_new)(const SSL_METHOD*);
  void (*SSL_CTX_free)(SSL_CTX*);
  int (*SSL_CTX_use_PrivateKey_file)(SSL_CTX*, const char*, int);
  int (*SSL_CTX_use_certificate_chain_file)(SSL_CTX*, const char*);
  int (*SSL_CTX_check_private_key)(const SSL_CTX*);
  int (*SSL_CTX_load_verify_locations)(SSL_CTX*, const char*, const char*);
  void (*SSL_CTX_set_verify)(SSL_CTX*, int, SSL_verify_cb);
  long (*SSL_CTX_ctrl)(SSL_CTX*, int, long, void*);
  void (*SSL_CTX_set_security_level)(SSL_CTX*, int);
  SSL* (*SSL_new)(SSL_CTX*);
  void (*SSL_free)(SSL*);
  int (*SSL_set_fd)(SSL*, int);
  void (*SSL_set_connect_state)(SSL*);
  void (*SSL_set_accept_state)(SSL*);
  int (*SSL_do_handshake)(SSL*);
  int (*SSL_read)(SSL*, void*, int);
  int (*SSL_write)(SSL*, const void*, int);
  int (*SSL_get_error)(const SSL*, int);
  int (*SSL_shutdown)(SSL*);
  int (*SSL_pending)(const SSL*);
  unsigned long (*ERR_get_error)();
  void (*ERR_clear_error)();
  void (*ERR_error_string_n)(unsigned long, char*, size_t);
};

// Throws InvalidOperationException when no usable libssl can be loaded.
const OpenSSL& openssl();
bool opensslAvailable();
std::string opensslLastError();

}  // namespace tls
}  // namespace tcp
}  // namespace transport
}  // namespace glb
