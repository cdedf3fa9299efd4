// transport::uv — source compatibility for code written against the reference's
// libuv transport (gloo/transport/uv/device.h:40-60). The reference uses libuv to get
// a portable (Windows / macOS) event loop under the same wire protocol as tcp; this
// build targets Linux only, so `uv::CreateDevice` hands out the epoll transport under
// the uv name. Nothing is lost: the reference's uv pairs support unbound buffers only
// (uv/pair.h:110-126 aborts on bound buffers and sync mode), the device returned here
// supports everything tcp does. libuv itself is neither needed nor linked.
#pragma once

#include <memory>
#include <string>

#include "glb/transport/device.h"

namespace glb {
namespace transport {
namespace uv {

struct attr {
  attr() = default;
  /* implicit */ attr(const char* host) : hostname(host) {}
  /* implicit */ attr(const std::string& host) : hostname(host) {}
  std::string hostname;
  std::string iface;
  int ai_family = 0;  // AF_UNSPEC
};

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr&);

}  // namespace uv
}  // namespace transport
}  // namespace glb
