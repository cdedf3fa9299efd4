// rendezvous::Context — a glb::Context that connects its full mesh through a
// Store; ContextFactory — mints additional contexts over an already connected
// one without touching the store again (address blobs travel over the backing
// context's own pairs). Parity: gloo/rendezvous/context.{h,cc}.
#pragma once

#include <memory>
#include <mutex>

#include "glb/context.h"
#include "glb/rendezvous/store.h"
#include "glb/transport/device.h"

namespace glb {
namespace rendezvous {

class Context : public ::glb::Context {
 public:
  Context(int rank, int size, int base = 2);
  ~Context() override;

  void connectFullMesh(std::shared_ptr<Store> store, std::shared_ptr<transport::Device>& dev);

 protected:
  friend class ContextFactory;
};

class ContextFactory {
 public:
  static constexpr auto kMaxAddressSize = 192;

  explicit ContextFactory(std::shared_ptr<::glb::Context> backingContext);

  // Collective over the backing context: every rank must call it, in the same order.
  std::shared_ptr<::glb::Context> makeContext(std::shared_ptr<transport::Device>& dev);

 protected:
  std::shared_ptr<::glb::Context> backingContext_;
  std::mutex mu_;
  uint32_t generation_ = 0;
};

}  // namespace rendezvous
}  // namespace glb
