// barrier (new-style): dissemination barrier, ceil(log2 P) rounds of zero-byte
// send/recv with partners at distance 2^k. Parity: gloo/barrier.{h,cc}.
#pragma once

#include "glb/collectives_common.h"

namespace glb {

class BarrierOptions : public detail::CollectiveOptionsBase {
 public:
  explicit BarrierOptions(const std::shared_ptr<Context>& context)
      : CollectiveOptionsBase(context), buffer(context->createUnboundBuffer(nullptr, 0)) {}
  std::unique_ptr<UnboundBuffer> buffer;
};

void barrier(BarrierOptions& opts);

}  // namespace glb
