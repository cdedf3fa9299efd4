// Shared plumbing for the new-style (options struct + free function) collectives:
// the options base class (context / tag / timeout), buffer holders that accept
// either a raw typed pointer or a ready-made UnboundBuffer, and range helpers.
#pragma once

#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "glb/common/logging.h"
#include "glb/context.h"
#include "glb/transport/unbound_buffer.h"
#include "glb/types.h"

namespace glb {

using transport::UnboundBuffer;

namespace detail {

class CollectiveOptionsBase {
 public:
  explicit CollectiveOptionsBase(const std::shared_ptr<Context>& context)
      : context(context), timeout(context->getTimeout()) {}

  void setTag(uint32_t t) { tag = t; }
  void setTimeout(std::chrono::milliseconds t) {
    GLB_ENFORCE(t.count() >= 0, "Invalid timeout ", t.count());
    timeout = t;
  }

  std::shared_ptr<Context> context;
  std::chrono::milliseconds timeout;
  uint32_t tag = 0;
};

// Half-open element range.
struct Range {
  size_t off = 0;
  size_t len = 0;
};

// i-th of `parts` near-equal pieces of r (the first r.len % parts pieces get one extra element).
// Largest payload (bytes) for which the latency-oriented one-hop variants are used:
// allreduce (no algorithm requested), allgather, broadcast, reduce. GLB_ONEHOP_MAX, default
// 16 KiB; 0 disables them.
size_t oneHopMaxBytes();

inline Range subRange(Range r, size_t parts, size_t i) {
  size_t base = r.len / parts;
  size_t rem = r.len % parts;
  Range out;
  out.off = r.off + i * base + std::min(i, rem);
  out.len = base + (i < rem ? 1 : 0);
  return out;
}

}  // namespace detail
}  // namespace glb
