#include "glb/common/trace.h"
#include "glb/gatherv.h"

namespace glb {

void gatherv(GathervOptions& opts) {
  GLB_HOST_TRACE("glb::gatherv");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.in != nullptr, "gatherv: input required");
  GLB_ENFORCE(opts.root >= 0 && opts.root < context->size, "gatherv: invalid root ", opts.root);
  GLB_ENFORCE(opts.elementSize > 0, "gatherv: element size not set");
  UnboundBuffer* in = opts.in.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kGatherSlotPrefix, opts.tag);
  if (r == opts.root) {
    GLB_ENFORCE(opts.out != nullptr, "gatherv: output required on root");
    GLB_ENFORCE_EQ(static_cast<int>(opts.elementsPerRank.size()), P, "gatherv: need one count per rank");
    UnboundBuffer* out = opts.out.get();
    std::vector<size_t> off(P + 1, 0);
    for (int i = 0; i < P; i++) off[i + 1] = off[i] + opts.elementsPerRank[i] * opts.elementSize;
    GLB_ENFORCE_GE(out->size, off[P], "gatherv: output too small");
    GLB_ENFORCE_EQ(in->size, off[r + 1] - off[r], "gatherv: root input size does not match its count");
    int posted = 0;
    for (int i = 0; i < P; i++) {
      if (i == r) continue;
      out->recv(i, slot, off[i], off[i + 1] - off[i]);
      posted++;
    }
    if (in->size > 0) std::memcpy(static_cast<char*>(out->ptr) + off[r], in->ptr, in->size);
    for (int i = 0; i < posted; i++) out->waitRecv(opts.timeout);
  } else {
    in->send(opts.root, slot);
    in->waitSend(opts.timeout);
  }
}

}  // namespace glb
