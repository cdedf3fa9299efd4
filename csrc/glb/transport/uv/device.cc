#include "glb/transport/uv/device.h"

#include "glb/transport/context.h"
#include "glb/transport/tcp/device.h"

namespace glb {
namespace transport {
namespace uv {

namespace {
// The "uv" device is the socket transport on the PORTABLE reactor: tcp::Loop with its poll(2)
// backend (pipe wake-up, no epoll / eventfd) - the role libuv plays in the reference
// (gloo/transport/uv/*: an event loop that exists on every platform). Pairs, buffers and the
// wire protocol are the tcp transport's, so a "uv" rank and a "tcp" rank interoperate.
class Device : public ::glb::transport::Device {
 public:
  explicit Device(std::shared_ptr<::glb::transport::Device> inner) : inner_(std::move(inner)) {}
  std::string str() const override { return "uv(" + inner_->str() + ")"; }
  const std::string& getPCIBusID() const override { return inner_->getPCIBusID(); }
  int getInterfaceSpeed() const override { return inner_->getInterfaceSpeed(); }
  bool hasGPUDirect() const override { return inner_->hasGPUDirect(); }
  std::shared_ptr<::glb::transport::Context> createContext(int rank, int size) override {
    return inner_->createContext(rank, size);
  }

 private:
  std::shared_ptr<::glb::transport::Device> inner_;
};
}  // namespace

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr& a) {
  tcp::attr t;
  t.hostname = a.hostname;
  t.iface = a.iface;
  t.ai_family = a.ai_family;
  t.portableLoop = true;
  return std::make_shared<Device>(tcp::CreateDevice(t));
}

}  // namespace uv
}  // namespace transport
}  // namespace glb
