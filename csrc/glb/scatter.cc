#include "glb/common/trace.h"
#include "glb/scatter.h"

namespace glb {

void scatter(ScatterOptions& opts) {
  GLB_HOST_TRACE("glb::scatter");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.out != nullptr, "scatter: output required");
  GLB_ENFORCE(opts.root >= 0 && opts.root < context->size, "scatter: invalid root ", opts.root);
  UnboundBuffer* out = opts.out.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kScatterSlotPrefix, opts.tag);
  if (r == opts.root) {
    GLB_ENFORCE_EQ(static_cast<int>(opts.in.size()), P, "scatter: root needs one input per rank");
    for (int i = 0; i < P; i++) GLB_ENFORCE_EQ(opts.in[i]->size, out->size, "scatter: input ", i, " size mismatch");
    for (int i = 0; i < P; i++) {
      if (i != r) opts.in[i]->send(i, slot);
    }
    if (out->size > 0) std::memcpy(out->ptr, opts.in[r]->ptr, out->size);
    for (int i = 0; i < P; i++) {
      if (i != r) opts.in[i]->waitSend(opts.timeout);
    }
  } else {
    out->recv(opts.root, slot);
    out->waitRecv(opts.timeout);
  }
}

}  // namespace glb
