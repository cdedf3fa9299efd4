#include "glb/rendezvous/context.h"

#include <cstring>

#include "glb/common/logging.h"
#include "glb/transport/context.h"
#include "glb/transport/unbound_buffer.h"
#include "glb/types.h"

namespace glb {
namespace rendezvous {

Context::Context(int rank, int size, int base) : ::glb::Context(rank, size, base) {}
Context::~Context() = default;

void Context::connectFullMesh(std::shared_ptr<Store> store, std::shared_ptr<transport::Device>& dev) {
  GLB_ENFORCE(store != nullptr, "connectFullMesh needs a store");
  GLB_ENFORCE(dev != nullptr, "connectFullMesh needs a device");
  auto transportContext = dev->createContext(rank, size);
  transportContext->setTimeout(getTimeout());
  transportContext->createAndConnectAllPairs(std::move(store));
  device_ = dev;
  transportContext_ = std::move(transportContext);
  transportContext_->onAttached(weak_from_this());
}

ContextFactory::ContextFactory(std::shared_ptr<::glb::Context> backingContext)
    : backingContext_(std::move(backingContext)) {
  GLB_ENFORCE(backingContext_ != nullptr);
}

// The new transport context exports one rendezvous blob; blobs are all-gathered
// over the backing context (sizes first, then payloads, each peer-to-peer), and
// the new context connects from them.
std::shared_ptr<::glb::Context> ContextFactory::makeContext(std::shared_ptr<transport::Device>& dev) {
  std::lock_guard<std::mutex> g(mu_);
  auto& back = backingContext_;
  const int rank = back->rank;
  const int size = back->size;

  auto context = std::shared_ptr<Context>(new Context(rank, size, back->base));
  context->setTimeout(back->getTimeout());
  auto tctx = dev->createContext(rank, size);
  tctx->setTimeout(back->getTimeout());

  std::vector<std::vector<char>> blobs(size);
  if (size > 1) {
    blobs[rank] = tctx->exportRendezvousBlob();
    const auto slot = Slot::build(kInternalSlotPrefix, generation_++);

    // Phase 1: exchange blob sizes.
    std::vector<uint64_t> sizes(size, 0);
    sizes[rank] = blobs[rank].size();
    {
      auto sendBuf = back->createUnboundBuffer(&sizes[rank], sizeof(uint64_t));
      std::vector<std::unique_ptr<transport::UnboundBuffer>> recvBufs;
      for (int i = 0; i < size; i++) {
        if (i == rank) continue;
        recvBufs.push_back(back->createUnboundBuffer(&sizes[i], sizeof(uint64_t)));
        recvBufs.back()->recv(i, slot);
      }
      for (int i = 0; i < size; i++) {
        if (i != rank) sendBuf->send(i, slot);
      }
      for (auto& b : recvBufs) b->waitRecv();
      for (int i = 0; i < size - 1; i++) sendBuf->waitSend();
    }
    // Phase 2: exchange blobs.
    {
      auto sendBuf = back->createUnboundBuffer(blobs[rank].data(), blobs[rank].size());
      std::vector<std::unique_ptr<transport::UnboundBuffer>> recvBufs;
      for (int i = 0; i < size; i++) {
        if (i == rank) continue;
        blobs[i].resize(sizes[i]);
        recvBufs.push_back(back->createUnboundBuffer(blobs[i].data(), blobs[i].size()));
        recvBufs.back()->recv(i, slot + 1);
      }
      for (int i = 0; i < size; i++) {
        if (i != rank) sendBuf->send(i, slot + 1);
      }
      for (auto& b : recvBufs) b->waitRecv();
      for (int i = 0; i < size - 1; i++) sendBuf->waitSend();
    }
    tctx->connectWithBlobs(blobs);
  }
  context->device_ = dev;
  context->transportContext_ = std::move(tctx);
  context->transportContext_->onAttached(std::weak_ptr<::glb::Context>(context));
  return context;
}

}  // namespace rendezvous
}  // namespace glb
