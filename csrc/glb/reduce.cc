#include "glb/common/trace.h"
#include "glb/reduce.h"

#include <cstring>
#include <vector>

namespace glb {

void reduce(ReduceOptions& opts) {
  GLB_HOST_TRACE("glb::reduce");
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
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kReduceSlotPrefix, opts.tag);
  // Small vectors: everyone sends straight to the root, which folds the contributions in
  // rank order - one hop instead of a ring reduce-scatter plus a gather.
  if (P > 1 && bytes <= detail::oneHopMaxBytes() && P <= 64) {
    UnboundBuffer* src = opts.in ? opts.in.get() : out;
    if (r != opts.root) {
      src->send(opts.root, slot, 0, bytes);
      src->waitSend(opts.timeout);
      return;
    }
    std::vector<char> tmpStorage(static_cast<size_t>(P) * bytes);
    auto tmp = context->createUnboundBuffer(tmpStorage.data(), tmpStorage.size());
    for (int q = 0; q < P; q++) {
      if (q != r) tmp->recv(q, slot, static_cast<size_t>(q) * bytes, bytes);
    }
    std::memcpy(tmpStorage.data() + static_cast<size_t>(r) * bytes, src->ptr, bytes);
    for (int q = 1; q < P; q++) tmp->waitRecv(opts.timeout);
    char* o = static_cast<char*>(out->ptr);
    opts.reduce(o, tmpStorage.data(), tmpStorage.data() + bytes, opts.elements);
    for (int q = 2; q < P; q++) opts.reduce(o, o, tmpStorage.data() + static_cast<size_t>(q) * bytes, opts.elements);
    return;
  }
  if (opts.in && opts.in->ptr != out->ptr) std::memcpy(out->ptr, opts.in->ptr, bytes);
  if (P == 1) return;

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
