// Host-side API of the CUDA-buffer collectives: argument checking, variant
// selection per message size and kernel launch. Every call is asynchronous on the
// given stream (a single kernel launch for registered buffers) and must be issued
// in the same order by every rank of the PeerContext.
//
// Buffers come in two flavours:
//   * PeerBuffer (registerBuffer / allocSymmetric): zero-copy, kernels read and
//     write the peers' copies directly;
//   * arbitrary device pointers: staged through the symmetric pool.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "glb/cuda/peer_context.h"
#include "glb/types.h"

namespace glb {
namespace cuda {

enum class AllreduceAlgo : int {
  AUTO = 0,
  ONE_SHOT = 1,
  TWO_SHOT = 2,
  NVLS = 3,
  // Literal schedules of the reference's named algorithms, executed by the
  // step-table kernel over peer pointers (schedule_kernels.cu).
  RING = 10,
  RING_CHUNKED = 11,
  HALVING_DOUBLING = 12,
  BCUBE = 13,
};

const char* allreduceAlgoName(AllreduceAlgo a);

struct Tuning {
  size_t oneShotMaxBytes = 256 * 1024;  // one-shot while bytes <= this / P; above: two-shot / NVLS
  size_t nvlsMinBytes = 32 * 1024;      // >= : NVLS when the buffer has a multicast alias (and P > 2)
  int maxBlocks = 64;                   // CTAs for the bandwidth kernels (clamped to co-residency cap)
  int oneShotBlocks = 8;
  bool nvlsReduceScatter = false;       // multimem.ld_reduce in reduce_scatter / reduce: measured slower than the
                                        // P2P pull at P=4 (403 vs 343 us @256 MB) and equal at P=8, so off by default
  int copyBlocks = 296;                 // CTAs for the store-only kernels (40 regs: 2-3 CTAs per SM)
  size_t bcastDirectMaxBytes = 256 * 1024;  // <= : root pushes everything itself
};
Tuning& tuning();
// Number of collective kernels launched by this process so far.
uint64_t launchCount();
void noteLaunch(unsigned n = 1);  // local (non-collective) kernels report here  // process-wide; initialised from GLB_CUDA_* env vars

// Which variant AUTO resolves to for this call.
AllreduceAlgo chooseAllreduce(const PeerContext& pc, size_t bytes, DataType dt, ReduceOp op, bool registered,
                              bool hasMulticast);

void barrier(PeerContext& pc, cudaStream_t stream);

// In place on a registered / symmetric buffer (count elements from byteOffset).
void allreduce(PeerContext& pc, const PeerBuffer& buf, size_t byteOffset, size_t count, DataType dt, ReduceOp op,
               AllreduceAlgo algo, cudaStream_t stream);

// Arbitrary device pointers (in may equal out); staged through the pool.
void allreduce(PeerContext& pc, const void* in, void* out, size_t count, DataType dt, ReduceOp op,
               AllreduceAlgo algo, cudaStream_t stream);


// ---- data movement -------------------------------------------------------------------
// `buf`/`out` arguments are registered or symmetric buffers; plain-pointer
// overloads stage through the pool (payload limited to the pool's bulk region).

// In place: root's bytes replace everyone's.
void broadcast(PeerContext& pc, const PeerBuffer& buf, size_t byteOffset, size_t bytes, int root,
               cudaStream_t stream);
void broadcast(PeerContext& pc, void* ptr, size_t bytes, int root, cudaStream_t stream);

// Rank r's `in` (bytesPerRank[r] bytes) lands at the prefix-sum offset in every out.
void allgatherv(PeerContext& pc, const void* in, const PeerBuffer& out, size_t outOffset,
                const std::vector<size_t>& bytesPerRank, cudaStream_t stream);
void allgatherv(PeerContext& pc, const void* in, void* out, const std::vector<size_t>& bytesPerRank,
                cudaStream_t stream);
// Only `root` receives.
void gatherv(PeerContext& pc, const void* in, const PeerBuffer& out, size_t outOffset,
             const std::vector<size_t>& bytesPerRank, int root, cudaStream_t stream);
void gatherv(PeerContext& pc, const void* in, void* out, const std::vector<size_t>& bytesPerRank, int root,
             cudaStream_t stream);

// Chunk j of my input (sendBytes[j] at the prefix-sum offset) goes to rank j, which
// stores what it gets from rank i at the prefix-sum offset of ITS recvBytes[i].
void alltoallv(PeerContext& pc, const void* in, const std::vector<size_t>& sendBytes, const PeerBuffer& out,
               size_t outOffset, const std::vector<size_t>& recvBytes, cudaStream_t stream);
void alltoallv(PeerContext& pc, const void* in, const std::vector<size_t>& sendBytes, void* out,
               const std::vector<size_t>& recvBytes, cudaStream_t stream);
// Root's input holds P chunks of `bytes`; chunk j lands in rank j's out.
void scatter(PeerContext& pc, const void* in, const PeerBuffer& out, size_t outOffset, size_t bytes, int root,
             cudaStream_t stream);
void scatter(PeerContext& pc, const void* in, void* out, size_t bytes, int root, cudaStream_t stream);

// Rank r gets elements [prefix(counts, r), +counts[r]) of the reduction in `out`.
void reduce_scatter(PeerContext& pc, const PeerBuffer& in, size_t inOffset, void* out,
                    const std::vector<size_t>& counts, DataType dt, ReduceOp op, cudaStream_t stream);
void reduce_scatter(PeerContext& pc, const void* in, void* out, const std::vector<size_t>& counts, DataType dt,
                    ReduceOp op, cudaStream_t stream);
// Reduction of every rank's `in` delivered to root's `out` (both registered).
void reduce(PeerContext& pc, const PeerBuffer& in, size_t inOffset, const PeerBuffer& out, size_t outOffset,
            size_t count, DataType dt, ReduceOp op, int root, cudaStream_t stream);
void reduce(PeerContext& pc, const void* in, void* out, size_t count, DataType dt, ReduceOp op, int root,
            cudaStream_t stream);

}  // namespace cuda
}  // namespace glb
