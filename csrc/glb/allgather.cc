#include "glb/common/trace.h"
#include "glb/allgather.h"

namespace glb {

void allgather(AllgatherOptions& opts) {
  GLB_HOST_TRACE("glb::allgather");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.out != nullptr, "allgather: output required");
  UnboundBuffer* in = opts.in.get();
  UnboundBuffer* out = opts.out.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kAllgatherSlotPrefix, opts.tag);
  GLB_ENFORCE_EQ(out->size % P, 0u, "allgather: output size must be a multiple of the context size");
  const size_t block = out->size / P;
  if (in != nullptr) {
    GLB_ENFORCE_EQ(in->size, block, "allgather: input size must equal output size / P");
    if (block > 0) std::memcpy(static_cast<char*>(out->ptr) + r * block, in->ptr, block);
  }
  if (P == 1 || block == 0) return;

  // Small blocks: one hop, everyone sends its block to everyone (P-1 tiny messages each)
  // instead of P-1 dependent ring steps.
  if (block <= detail::oneHopMaxBytes() && P <= 32) {
    for (int k = 1; k < P; k++) {
      const int q = (r - k + P) % P;
      out->recv(q, slot, static_cast<size_t>(q) * block, block);
    }
    for (int k = 1; k < P; k++) out->send((r + k) % P, slot, static_cast<size_t>(r) * block, block);
    for (int k = 1; k < P; k++) out->waitRecv(opts.timeout);
    for (int k = 1; k < P; k++) out->waitSend(opts.timeout);
    return;
  }

  const int right = (r + 1) % P;
  const int left = (r - 1 + P) % P;
  // Two half-blocks per step keep the pipe full: while half A of step s is still
  // arriving, half B of step s-1 is already being forwarded.
  const size_t h0 = block / 2;
  const size_t h1 = block - h0;
  for (int s = 0; s < P - 1; s++) {
    const size_t sendOff = static_cast<size_t>((r - s + P) % P) * block;
    const size_t recvOff = static_cast<size_t>((r - s - 1 + 2 * P) % P) * block;
    out->recv(left, slot, recvOff, h0);
    out->recv(left, slot, recvOff + h0, h1);
    out->send(right, slot, sendOff, h0);
    out->send(right, slot, sendOff + h0, h1);
    out->waitRecv(opts.timeout);
    out->waitRecv(opts.timeout);
    out->waitSend(opts.timeout);
    out->waitSend(opts.timeout);
  }
}

}  // namespace glb
