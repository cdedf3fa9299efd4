#include "glb/reduce.h"

namespace glb {

void reduce(ReduceOptions& opts) {
  const auto& context = opts.context;
  GLB_ENFORCE(opts.out != nullptr, "reduce: output required (used as scratch on non-root ranks)");
  GLB_ENFORCE(opts.root >= 0 && opts.root < context->size, "reduce: invalid root ", opts.root);
  GLB_ENFORCE(opts.elementSize > 0, "reduce: element size not set");
  GLB_ENFORCE(static_cast<bool>(opts.reduce), "reduce: reduce function not set");
  UnboundBuffer* out = opts.out.get();
  const size_t bytes = opts.elements * opts.elementSize;
  GLB_ENFORCE_EQ(out->size, bytes, "reduce: output size mismatch");
  if (opts.in) GLB_ENFORCE_EQ(opts.in->size, bytes, "reduce: input size mismatch");
  if (opts.elements == 0) return;
  if (opts.in && opts.in->ptr != out->ptr) std::memcpy(out->ptr, opts.in->ptr, bytes);
  const int P = context->size;
  const int r = context->rank;
  if (P == 1) return;

  const auto slot = Slot::build(kReduceSlotPrefix, opts.tag);
  detail::ringReduceScatter(context, out, opts.elements, opts.elementSize, opts.reduce, opts.maxSegmentSize,
                            slot, opts.timeout);
  // Rank i now owns the reduced chunk (i + 1) % P; ship it to the root.
  const detail::Range all{0, opts.elements};
  if (r == opts.root) {
    int posted = 0;
    for (int i = 0; i < P; i++) {
      if (i == r) continue;
      auto c = detail::subRange(all, P, (i + 1) % P);
      out->recv(i, slot + 1, c.off * opts.elementSize, c.len * opts.elementSize);
      posted++;
    }
    for (int i = 0; i < posted; i++) out->waitRecv(opts.timeout);
  } else {
    auto c = detail::subRange(all, P, (r + 1) % P);
    out->send(opts.root, slot + 1, c.off * opts.elementSize, c.len * opts.elementSize);
    out->waitSend(opts.timeout);
  }
}

}  // namespace glb
