// Literal schedules of the reference's named allreduce algorithms (ring,
// ring_chunked, halving_doubling, bcube), expressed as step tables that ONE generic
// kernel executes over peer pointers (schedule_kernels.cu). Every step pulls a range
// from a set of peers and either folds it into the local buffer or overwrites it;
// steps are separated by device-side barriers, so the whole algorithm — all 2(P-1)
// ring rounds, or 2·lg P halving-doubling steps — is a single kernel launch.
// These exist so that "cuda_allreduce_halving_doubling at 2/4/8 GPUs" is a real,
// measurable variant; AUTO normally picks one-shot / two-shot / NVLS instead.
// Cost models: gloo docs/algorithms.md:33-129.
#pragma once

#include <vector>

#include "glb/cuda/collectives.h"
#include "glb/cuda/comm_types.h"

namespace glb {
namespace cuda {

constexpr int kSchedMaxPeers = kMaxRanks - 1;

enum SchedMode : int {
  SCHED_REDUCE = 0,  // local[range] = local[range] (op) peer_0[range] (op) peer_1[range] ...
  SCHED_COPY = 1,    // local[range] = peer_0[range]
  SCHED_STAGE = 2,   // pool[range] = local[range]   (no peer traffic)
};

struct SchedStep {
  int mode;
  int npeers;
  int peers[kSchedMaxPeers];
  int fromStage;            // 1: read the peers' pool copy instead of their user buffer
  int sync;                 // 1: a device barrier separates this step from the previous one;
                            // 0: it belongs to the same phase (pipelined schedules run the
                            // steps of several chunks between two barriers)
  unsigned long long off;   // element offset
  unsigned long long len;   // element count
};

struct Schedule {
  std::vector<SchedStep> steps;  // executed in order, one barrier before each + one at the end
  const char* name = "";
  bool needsStage = false;       // uses SCHED_STAGE: message must fit the pool's bulk region
};

// Builders: the table for `rank` out of `size`, over `count` elements.
// `packElems` = elements per 16-byte pack (ranges are aligned to it).
Schedule buildRingSchedule(int rank, int size, size_t count, size_t packElems);
Schedule buildRingChunkedSchedule(int rank, int size, size_t count, size_t packElems);
Schedule buildHalvingDoublingSchedule(int rank, int size, size_t count, size_t packElems);
Schedule buildBcubeSchedule(int rank, int size, size_t count, int base, size_t packElems);
// Halving-doubling over `chunks` sub-vectors, skewed by one step each: while chunk c runs
// step s, chunk c+1 runs step s-1 in the same barrier phase (the reference's
// CudaAllreduceHalvingDoublingPipelined, cuda_allreduce_halving_doubling.cc:412-455,
// overlaps the local reduction of one chunk with the transfer of the next the same way).
Schedule buildHalvingDoublingPipelinedSchedule(int rank, int size, size_t count, size_t packElems, int chunks = 2);
// Number of device barriers a schedule needs (steps with sync = 1, plus the closing one).
int scheduleBarriers(const Schedule& s);

// kernels.h-style launcher (defined in schedule_kernels.cu). `table` is device memory.
void launchSchedule(const CommArgs& a, const PeerPtrs& bufs, const PeerPtrs& stage, const SchedStep* table,
                    int nsteps, int nbarriers, DataType dt, ReduceOp op, float scale, size_t count, bool vectorOk,
                    int blocks, cudaStream_t stream);

}  // namespace cuda
}  // namespace glb
