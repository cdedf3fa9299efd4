#include "glb/common/trace.h"
#include "glb/barrier.h"

namespace glb {

void barrier(BarrierOptions& opts) {
  GLB_HOST_TRACE("glb::barrier");
  const auto& context = opts.context;
  auto& buffer = opts.buffer;
  const auto slot = Slot::build(kBarrierSlotPrefix, opts.tag);
  const int P = context->size;
  const int r = context->rank;
  // After round k every rank has (transitively) heard from 2^(k+1) ranks.
  int round = 0;
  for (int d = 1; d < P; d <<= 1, round++) {
    buffer->recv((r - d + P) % P, slot + round);
    buffer->send((r + d) % P, slot + round);
    buffer->waitRecv(opts.timeout);
    buffer->waitSend(opts.timeout);
  }
}

}  // namespace glb
