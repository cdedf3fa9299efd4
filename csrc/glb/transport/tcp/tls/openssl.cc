#include "glb/transport/tcp/tls/openssl.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

#include "glb/common/logging.h"

namespace glb {
namespace transport {
namespace tcp {
namespace tls {

namespace {
OpenSSL gApi;
bool gLoaded = false;
std::string gError;
std::once_flag gOnce;

void load() {
  void* ssl = nullptr;
  for (const char* name : {"libssl.so.3", "libssl.so.1.1", "libssl.so"}) {
    ssl = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (ssl != nullptr) break;
  }
  void* crypto = nullptr;
  for (const char* name : {"libcrypto.so.3", "libcrypto.so.1.1", "libcrypto.so"}) {
    crypto = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (crypto != nullptr) break;
  }
  if (ssl == nullptr || crypto == nullptr) {
    gError = "libssl / libcrypto not found";
    return;
  }
  bool ok = true;
  auto sym = [&](void* lib, const char* n) {
    void* p = dlsym(lib, n);
    if (p == nullptr) {
      ok = false;
      gError = std::string("missing OpenSSL symbol ") + n;
    }
    return p;
  };
#define GLB_SSL(lib, name) gApi.name = reinterpret_cast<decltype(gApi.name)>(sym(lib, #name))
  GLB_SSL(ssl, TLS_method);
  GLB_SSL(ssl, SSL_CTX_new);
  GLB_SSL(ssl, SSL_CTX_free);
  GLB_SSL(ssl, SSL_CTX_use_PrivateKey_file);
  GLB_SSL(ssl, SSL_CTX_use_certificate_chain_file);
  GLB_SSL(ssl, SSL_CTX_check_private_key);
  GLB_SSL(ssl, SSL_CTX_load_verify_locations);
  GLB_SSL(ssl, SSL_CTX_set_verify);
  GLB_SSL(ssl, SSL_CTX_ctrl);
  GLB_SSL(ssl, SSL_CTX_set_security_level);
  GLB_SSL(ssl, SSL_new);
  GLB_SSL(ssl, SSL_free);
  GLB_SSL(ssl, SSL_set_fd);
  GLB_SSL(ssl, SSL_set_connect_state);
  GLB_SSL(ssl, SSL_set_accept_state);
  GLB_SSL(ssl, SSL_do_handshake);
  GLB_SSL(ssl, SSL_read);
  GLB_SSL(ssl, SSL_write);
  GLB_SSL(ssl, SSL_get_error);
  GLB_SSL(ssl, SSL_shutdown);
  GLB_SSL(ssl, SSL_pending);
  GLB_SSL(crypto, ERR_get_error);
  GLB_SSL(crypto, ERR_clear_error);
  GLB_SSL(crypto, ERR_error_string_n);
#undef GLB_SSL
  gLoaded = ok;
}
}  // namespace

bool opensslAvailable() {
  std::call_once(gOnce, load);
  return gLoaded;
}

const OpenSSL& openssl() {
  if (!opensslAvailable()) GLB_THROW_INVALID_OPERATION_EXCEPTION("OpenSSL unavailable: ", gError);
  return gApi;
}

std::string opensslLastError() {
  if (!opensslAvailable()) return gError;
  std::string out;
  while (unsigned long e = gApi.ERR_get_error()) {
    char buf[256];
    gApi.ERR_error_string_n(e, buf, sizeof(buf));
    if (!out.empty()) out += "; ";
    out += buf;
  }
  return out.empty() ? "unknown TLS error" : out;
}

}  // namespace tls
}  // namespace tcp
}  // namespace transport
}  // namespace glb
