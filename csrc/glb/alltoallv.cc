#include "glb/common/trace.h"
#include "glb/alltoallv.h"

namespace glb {

void alltoallv(AlltoallvOptions& opts) {
  GLB_HOST_TRACE("glb::alltoallv");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.in != nullptr && opts.out != nullptr, "alltoallv: input and output required");
  UnboundBuffer* in = opts.in.get();
  UnboundBuffer* out = opts.out.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kAlltoallSlotPrefix, opts.tag);
  GLB_ENFORCE_EQ(static_cast<int>(opts.inLengthPerRank.size()), P);
  GLB_ENFORCE_EQ(static_cast<int>(opts.outLengthPerRank.size()), P);
  GLB_ENFORCE_GE(in->size, opts.inOffsetPerRank[P - 1] + opts.inLengthPerRank[P - 1], "alltoallv: input too small");
  GLB_ENFORCE_GE(out->size, opts.outOffsetPerRank[P - 1] + opts.outLengthPerRank[P - 1], "alltoallv: output too small");
  GLB_ENFORCE_EQ(opts.inLengthPerRank[r], opts.outLengthPerRank[r], "alltoallv: self chunk size differs between input and output");
  if (opts.inLengthPerRank[r] > 0) {
    std::memcpy(static_cast<char*>(out->ptr) + opts.outOffsetPerRank[r],
                static_cast<char*>(in->ptr) + opts.inOffsetPerRank[r], opts.inLengthPerRank[r]);
  }
  if (P == 1) return;
  for (int i = 1; i < P; i++) {
    const int src = (r - i + P) % P;
    out->recv(src, slot, opts.outOffsetPerRank[src], opts.outLengthPerRank[src]);
  }
  for (int i = 1; i < P; i++) {
    const int dst = (r + i) % P;
    in->send(dst, slot, opts.inOffsetPerRank[dst], opts.inLengthPerRank[dst]);
  }
  for (int i = 1; i < P; i++) {
    out->waitRecv(opts.timeout);
    in->waitSend(opts.timeout);
  }
}

}  // namespace glb
