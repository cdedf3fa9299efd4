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
constexpr int kP2pLanes = 64;     // max CTAs per direction of a point-to-point transfer

// Why a kernel gave up (SignalPad::abort / the host status word).
enum AbortCode : uint32_t {
  kAbortNone = 0,
  kAbortTimeout = 1,   // a peer did not reach the barrier within the timeout
  kAbortPeer = 2,      // a peer (or another CTA) aborted first and told us
};

// One per rank, at the start of its symmetric pool. `flag[b][s]` is written by
// rank s's block b; everything else is only touched by the owning rank (except
// `abort`, which any rank may raise).
struct alignas(128) SignalPad {
  uint32_t flag[kMaxBlocks][kMaxRanks];
  uint32_t epoch;      // barrier epoch consumed so far (advanced by the last CTA to finish)
  uint32_t stageSeq;   // number of staged (one-shot) launches so far -> double-buffer parity
  uint32_t done;       // CTA completion ticket
  uint32_t llSeq;      // number of flag-in-data (LL) launches so far -> flag value / parity
  uint32_t abort;      // != 0: a barrier timed out somewhere; every later barrier returns at once
  uint32_t abortRank;  // who was missing when the first local time-out fired
  uint32_t pad[26];
  // Scratch for in-kernel metadata exchange (alltoallv receive offsets).
  unsigned long long xchg[kMaxRanks];
  // Point-to-point lanes (p2p_kernels.cu). Lane l of the pair (s -> d) counts mailbox chunks:
  //   d.pad.p2pHead[l][s]  written by s: chunks published so far
  //   s.pad.p2pTail[l][d]  written by d: chunks consumed so far
  //   p2pSent / p2pRecvd   the owners' own running totals (persist across launches)
  uint32_t p2pHead[kP2pLanes][kMaxRanks];
  uint32_t p2pTail[kP2pLanes][kMaxRanks];
  uint32_t p2pSent[kP2pLanes][kMaxRanks];
  uint32_t p2pRecvd[kP2pLanes][kMaxRanks];
};

struct CommArgs {
  int rank;
  int nranks;
  SignalPad* sig[kMaxRanks];  // sig[r] = rank r's pad as mapped in this process
  SignalPad* self;            // == sig[rank]; a plain field so that kernels need no dynamic
                              // index into their parameter space to reach their own pad
  // Failure detection: a barrier that waits longer than this sets the abort word,
  // reports through `hostStatus` (mapped pinned host memory) and lets the kernel exit.
  unsigned long long timeoutNs;  // 0 = wait forever
  uint32_t* hostStatus;          // may be null
};

struct PeerPtrs {
  void* p[kMaxRanks];
};

// Extra local pointers of a multi-pointer call (beyond the one that takes part in the
// exchange): folded into it before, and overwritten with the result after, inside the
// same kernel.
constexpr int kMaxLocal = 7;
struct LocalPtrs {
  int n = 0;
  void* p[kMaxLocal] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// How a bandwidth kernel is launched: CTAs and the per-thread unroll (independent
// 128-bit accesses in flight). Filled from the tuning table (tuning.h).
struct LaunchCfg {
  int blocks = 64;
  int unroll = 0;  // 0 = the kernel's default for this P
};

}  // namespace cuda
}  // namespace glb
