#include "glb/common/trace.h"
#include "glb/gather.h"

namespace glb {

void gather(GatherOptions& opts) {
  GLB_HOST_TRACE("glb::gather");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.in != nullptr, "gather: input required");
  GLB_ENFORCE(opts.root >= 0 && opts.root < context->size, "gather: invalid root ", opts.root);
  UnboundBuffer* in = opts.in.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kGatherSlotPrefix, opts.tag);
  if (r == opts.root) {
    GLB_ENFORCE(opts.out != nullptr, "gather: output required on root");
    UnboundBuffer* out = opts.out.get();
    const size_t chunk = in->size;
    GLB_ENFORCE_EQ(out->size, chunk * P, "gather: output must hold P inputs");
    for (int i = 0; i < P; i++) {
      if (i != r) out->recv(i, slot, i * chunk, chunk);
    }
    if (chunk > 0) std::memcpy(static_cast<char*>(out->ptr) + r * chunk, in->ptr, chunk);
    for (int i = 0; i < P - 1; i++) out->waitRecv(opts.timeout);
  } else {
    in->send(opts.root, slot);
    in->waitSend(opts.timeout);
  }
}

}  // namespace glb
