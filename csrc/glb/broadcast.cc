#include "glb/common/trace.h"
#include "glb/broadcast.h"

#include <algorithm>

#include "glb/common/utils.h"

namespace glb {

void broadcast(BroadcastOptions& opts) {
  GLB_HOST_TRACE("glb::broadcast");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.out != nullptr, "broadcast: output required");
  GLB_ENFORCE(opts.root >= 0 && opts.root < context->size, "broadcast: invalid root ", opts.root);
  UnboundBuffer* out = opts.out.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kBroadcastSlotPrefix, opts.tag);
  if (r == opts.root && opts.in) {
    GLB_ENFORCE_EQ(opts.in->size, out->size, "broadcast: input/output size mismatch");
    if (opts.in->ptr != out->ptr && out->size > 0) std::memcpy(out->ptr, opts.in->ptr, out->size);
  }
  if (P == 1) return;

  // Small payloads: the root sends to everyone itself. P-1 back-to-back writes cost a few
  // microseconds; every level of the tree costs a full message hop.
  if (out->size <= detail::oneHopMaxBytes() && P <= 32) {
    if (r == opts.root) {
      for (int k = 1; k < P; k++) out->send((r + k) % P, slot, 0, out->size);
      for (int k = 1; k < P; k++) out->waitSend(opts.timeout);
    } else {
      out->recv(opts.root, slot, 0, out->size);
      out->waitRecv(opts.timeout);
    }
    return;
  }

  const int vrank = (r - opts.root + P) % P;
  // Parent: clear the lowest set bit of vrank. Children: vrank + 2^k for 2^k below that bit.
  int lowbit = 1;
  while (lowbit < P && (vrank & lowbit) == 0) lowbit <<= 1;  // for root ends >= P
  const int parent = vrank == 0 ? -1 : ((vrank & ~lowbit) + opts.root) % P;
  std::vector<int> children;
  for (int d = (vrank == 0 ? (1 << log2ceil(static_cast<uint32_t>(P))) : lowbit) >> 1; d >= 1; d >>= 1) {
    if (vrank + d < P) children.push_back((vrank + d + opts.root) % P);
  }

  const size_t bytes = out->size;
  const size_t seg = std::max<size_t>(1, opts.maxSegmentSize);
  const size_t nseg = std::max<size_t>(1, ceilDiv(bytes, seg));
  if (parent >= 0) {
    for (size_t j = 0; j < nseg; j++) {
      size_t off = j * seg;
      out->recv(parent, slot, off, std::min(seg, bytes - std::min(bytes, off)));
    }
  }
  size_t sends = 0;
  for (size_t j = 0; j < nseg; j++) {
    if (parent >= 0) out->waitRecv(opts.timeout);
    size_t off = j * seg;
    size_t len = std::min(seg, bytes - std::min(bytes, off));
    for (int c : children) {
      out->send(c, slot, off, len);
      sends++;
    }
  }
  for (size_t k = 0; k < sends; k++) out->waitSend(opts.timeout);
}

}  // namespace glb
