// Thin NCCL wrapper, loaded at run time (dlopen libnccl.so.2 — the copy torch already
// loaded when running under PyTorch). It exists as the *baseline comparator* for the
// fused kernels and for parity with the reference's gloo::nccl ops (reduce, allreduce,
// reduce_scatter, broadcast, allgather; nccl/nccl.{h,cu}); no collective in this
// library is implemented by calling NCCL.
//   NcclComm::initRank   one communicator rank per process (unique id travels over the
//                        glb::Context)
//   NcclComm::initAll    one communicator per local device list inside a process
//                        (the reference's ncclCommInitAll usage)
#pragma once

#include <cuda_runtime.h>

#include <memory>
#include <vector>

#include "glb/context.h"
#include "glb/types.h"

namespace glb {
namespace cuda {

bool ncclAvailable();
std::string ncclVersionString();

class NcclComm {
 public:
  static std::shared_ptr<NcclComm> initRank(const std::shared_ptr<Context>& ctx, int device);
  static std::vector<std::shared_ptr<NcclComm>> initAll(const std::vector<int>& devices);
  ~NcclComm();

  int rank() const { return rank_; }
  int size() const { return size_; }
  int device() const { return device_; }

  void allreduce(const void* src, void* dst, size_t count, DataType dt, ReduceOp op, cudaStream_t stream);
  void reduce(const void* src, void* dst, size_t count, DataType dt, ReduceOp op, int root, cudaStream_t stream);
  void reduceScatter(const void* src, void* dst, size_t recvCount, DataType dt, ReduceOp op, cudaStream_t stream);
  void broadcast(const void* src, void* dst, size_t count, DataType dt, int root, cudaStream_t stream);
  void allgather(const void* src, void* dst, size_t sendCount, DataType dt, cudaStream_t stream);
  // alltoall through grouped send/recv (what NCCL users write by hand).
  void alltoall(const void* src, void* dst, size_t countPerRank, DataType dt, cudaStream_t stream);

  // Grouped ncclSend + ncclRecv: the neighbour exchange of ring attention / pipeline stages.
  void sendrecv(const void* src, int dst, void* dstBuf, int srcRank, size_t count, DataType dt, cudaStream_t stream);
  // NCCL-owned, communicator-registered buffers (ncclMemAlloc + ncclCommRegister).
  void* memAlloc(size_t bytes);
  void memFree(void* p);
  void* registerBuffer(void* p, size_t bytes);  // returns the registration handle
  void deregisterBuffer(void* handle);

  static void groupStart();
  static void groupEnd();

 private:
  NcclComm() = default;
  void* comm_ = nullptr;
  int rank_ = 0, size_ = 1, device_ = 0;
};

}  // namespace cuda
}  // namespace glb
