#include "glb/common/trace.h"
#include "glb/alltoall.h"

namespace glb {

void alltoall(AlltoallOptions& opts) {
  GLB_HOST_TRACE("glb::alltoall");
  const auto& context = opts.context;
  GLB_ENFORCE(opts.in != nullptr && opts.out != nullptr, "alltoall: input and output required");
  UnboundBuffer* in = opts.in.get();
  UnboundBuffer* out = opts.out.get();
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kAlltoallSlotPrefix, opts.tag);
  GLB_ENFORCE_EQ(in->size, out->size, "alltoall: input and output sizes differ");
  GLB_ENFORCE_EQ(in->size % P, 0u, "alltoall: buffer size must be a multiple of the context size");
  const size_t chunk = in->size / P;
  if (chunk > 0) {
    std::memcpy(static_cast<char*>(out->ptr) + r * chunk, static_cast<char*>(in->ptr) + r * chunk, chunk);
  }
  if (P == 1) return;
  for (int i = 1; i < P; i++) {
    const int src = (r - i + P) % P;
    out->recv(src, slot, src * chunk, chunk);
  }
  for (int i = 1; i < P; i++) {
    const int dst = (r + i) % P;
    in->send(dst, slot, dst * chunk, chunk);
  }
  for (int i = 1; i < P; i++) {
    out->waitRecv(opts.timeout);
    in->waitSend(opts.timeout);
  }
}

}  // namespace glb
