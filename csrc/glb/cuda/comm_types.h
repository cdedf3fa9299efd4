// Plain structs shared between host code and device kernels (safe to include
// from .cc files): the per-rank signal pad and the by-value kernel arguments.
#pragma once

#include <cstddef>
#include <cstdint>

namespace glb {
namespace cuda {

constexpr int kMaxRanks = 16;
constexpr int kMaxBlocks = 512;   // upper bound on CTAs of one collective kernel
constexpr int kThreads = 512;     // threads per CTA of the data-moving kernels

// One per rank, at the start of its symmetric pool. `flag[b][s]` is written by
// rank s's block b; everything else is only touched by the owning rank.
struct alignas(128) SignalPad {
  uint32_t flag[kMaxBlocks][kMaxRanks];
  uint32_t epoch;      // barrier epoch consumed so far (advanced by the last CTA to finish)
  uint32_t stageSeq;   // number of staged (one-shot) launches so far -> double-buffer parity
  uint32_t done;       // CTA completion ticket
  uint32_t pad[29];
  // Scratch for in-kernel metadata exchange (alltoallv receive offsets).
  unsigned long long xchg[kMaxRanks];
};

struct CommArgs {
  int rank;
  int nranks;
  SignalPad* sig[kMaxRanks];  // sig[r] = rank r's pad as mapped in this process
};

struct PeerPtrs {
  void* p[kMaxRanks];
};

}  // namespace cuda
}  // namespace glb
