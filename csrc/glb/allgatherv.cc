#include "glb/common/trace.h"
#include "glb/allgatherv.h"

namespace glb {

void allgatherv(AllgathervOptions& opts) {
  GLB_HOST_TRACE("glb::allgatherv");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.out != nullptr, "allgatherv: output required");
  GLB_ENFORCE(opts.elementSize > 0, "allgatherv: element size not set");
  const int P = context->size;
  const int r = context->rank;
  GLB_ENFORCE_EQ(static_cast<int>(opts.elements.size()), P, "allgatherv: need one count per rank");
  UnboundBuffer* in = opts.in.get();
  UnboundBuffer* out = opts.out.get();
  const auto slot = Slot::build(kAllgatherSlotPrefix, opts.tag);

  std::vector<size_t> off(P + 1, 0);
  for (int i = 0; i < P; i++) off[i + 1] = off[i] + opts.elements[i] * opts.elementSize;
  GLB_ENFORCE_GE(out->size, off[P], "allgatherv: output too small for the sum of counts");
  const size_t mine = off[r + 1] - off[r];
  if (in != nullptr) {
    GLB_ENFORCE_GE(in->size, mine, "allgatherv: input smaller than this rank's count");
    if (mine > 0) std::memcpy(static_cast<char*>(out->ptr) + off[r], in->ptr, mine);
  }
  if (P == 1) return;

  const int right = (r + 1) % P;
  const int left = (r - 1 + P) % P;
  for (int s = 0; s < P - 1; s++) {
    const int sendIdx = (r - s + P) % P;
    const int recvIdx = (r - s - 1 + 2 * P) % P;
    // Zero-length blocks still travel (as empty messages) so the ring stays in step.
    out->recv(left, slot, off[recvIdx], off[recvIdx + 1] - off[recvIdx]);
    out->send(right, slot, off[sendIdx], off[sendIdx + 1] - off[sendIdx]);
    out->waitRecv(opts.timeout);
    out->waitSend(opts.timeout);
  }
}

}  // namespace glb
