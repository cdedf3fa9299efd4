#include "glb/common/trace.h"
#include "glb/reduce_scatter.h"

namespace glb {

void reduce_scatter(ReduceScatterOptions& opts) {
  GLB_HOST_TRACE("glb::reduce_scatter");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.in != nullptr && opts.out != nullptr, "reduce_scatter: input and output required");
  GLB_ENFORCE(opts.elementSize > 0, "reduce_scatter: element size not set");
  GLB_ENFORCE(static_cast<bool>(opts.reduce), "reduce_scatter: reduce function not set");
  const int P = context->size;
  const int r = context->rank;
  const size_t es = opts.elementSize;
  UnboundBuffer* in = opts.in.get();
  UnboundBuffer* out = opts.out.get();
  const size_t total = in->size / es;

  std::vector<size_t> counts = opts.recvCounts;
  if (counts.empty()) {
    for (int i = 0; i < P; i++) counts.push_back(detail::subRange({0, total}, P, i).len);
  }
  GLB_ENFORCE_EQ(static_cast<int>(counts.size()), P, "reduce_scatter: need one recv count per rank");
  std::vector<size_t> off(P + 1, 0);
  for (int i = 0; i < P; i++) off[i + 1] = off[i] + counts[i];
  GLB_ENFORCE_EQ(off[P], total, "reduce_scatter: recv counts must add up to the input length");
  GLB_ENFORCE_GE(out->size, counts[r] * es, "reduce_scatter: output too small");
  const size_t mine = counts[r];

  if (mine > 0) std::memcpy(out->ptr, static_cast<char*>(in->ptr) + off[r] * es, mine * es);
  if (P == 1) return;
  const auto slot = Slot::build(kReduceScatterSlotPrefix, opts.tag);

  std::vector<char> tmpStorage(std::max<size_t>(1, (P - 1) * mine * es));
  auto tmp = context->createUnboundBuffer(tmpStorage.data(), tmpStorage.size());
  for (int i = 1; i < P; i++) {
    const int src = (r - i + P) % P;
    tmp->recv(src, slot, (i - 1) * mine * es, mine * es);
  }
  for (int i = 1; i < P; i++) {
    const int dst = (r + i) % P;
    in->send(dst, slot, off[dst] * es, counts[dst] * es);
  }
  for (int n = 1; n < P; n++) {
    int src = -1;
    tmp->waitRecv(&src, opts.timeout);
    const int i = (r - src + P) % P;
    if (mine > 0) opts.reduce(out->ptr, out->ptr, tmpStorage.data() + (i - 1) * mine * es, mine);
  }
  for (int n = 1; n < P; n++) in->waitSend(opts.timeout);
}

}  // namespace glb
