// Local (intra-process) building blocks of the CUDA algorithms, as LocalOp objects with the
// reference's asynchronous contract: runAsync() enqueues on the op's stream(s), wait() blocks
// the host until the result is usable. Parity:
//   CudaLocalMemcpy                       gloo/cuda.h:247-272
//   CudaLocalNativeReduce / Broadcast     gloo/cuda_collectives_native.h:61-130, 210-257
//   CudaLocalHostReduce / Broadcast       gloo/cuda_collectives_host.h:22-136, 143-214
//   CudaLocalNCCLReduce / Broadcast       gloo/cuda_collectives_nccl.h:40-229
//   cudaDevice{Reduce,Broadcast}, cudaHost{Reduce,Broadcast}   gloo/cuda_collectives_device.h:28-75,
//                                                             cuda_collectives_host.h:216-260
//   findCudaDevicePointerClosestToDevice  gloo/cuda_private.h:64-100
// Differences: the native reduce is ONE kernel over all sources (the reference reduces
// pairwise in a tree, one launch per pair); sources may live on several GPUs of the process
// (read through peer access); fp16 / bf16 accumulate in fp32.
#pragma once

#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/cuda/algorithms.h"
#include "glb/cuda/nccl_wrapper.h"
#include "glb/cuda/stream.h"
#include "glb/transport/device.h"

namespace glb {

using cuda::CudaDevicePointer;
using cuda::CudaHostPointer;
using cuda::CudaStream;

// Asynchronous copy between any two of {device pointer, pinned host pointer}.
template <typename T, typename Src, typename Dst>
class CudaLocalMemcpy : public LocalOp<T> {
 public:
  CudaLocalMemcpy(CudaStream& stream, Src& src, Dst& dst, size_t offset, size_t count)
      : stream_(stream), src_(src.range(offset, count)), dst_(dst.range(offset, count)) {}
  void runAsync() override { stream_.copyAsync(*dst_, *src_, src_.getCount() * sizeof(T)); }
  void wait() override { stream_.wait(); }

 private:
  CudaStream& stream_;
  Src src_;
  Dst dst_;
};

// Reduce N device buffers into `target` (a device pointer or a pinned host pointer).
template <typename T, typename Dst>
class CudaLocalNativeReduce : public LocalOp<T> {
 public:
  CudaLocalNativeReduce(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs, Dst& target,
                        const CudaReductionFunction<T>* fn, size_t offset, size_t count);
  void runAsync() override;
  void wait() override;

 private:
  std::vector<CudaStream>& streams_;
  std::vector<CudaDevicePointer<T>> srcs_;
  Dst target_;
  const CudaReductionFunction<T>* fn_;
  bool targetIsFirst_ = false;
};

// Copy `source` (device or pinned host) to N device buffers.
template <typename T, typename Src>
class CudaLocalNativeBroadcast : public LocalOp<T> {
 public:
  CudaLocalNativeBroadcast(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs, Src& source,
                           size_t offset, size_t count);
  void runAsync() override;
  void wait() override;

 private:
  std::vector<CudaStream>& streams_;
  std::vector<CudaDevicePointer<T>> dsts_;
  Src source_;
};

// Reduce through pinned host memory: D2H every buffer, reduce on the CPU.
template <typename T>
class CudaLocalHostReduce : public LocalOp<T> {
 public:
  CudaLocalHostReduce(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                      CudaHostPointer<T>& target, const CudaReductionFunction<T>* fn, size_t offset, size_t count);
  void runAsync() override;
  void wait() override;

 private:
  std::vector<CudaStream>& streams_;
  std::vector<CudaDevicePointer<T>> srcs_;
  CudaHostPointer<T> target_;
  const CudaReductionFunction<T>* fn_;
  std::vector<CudaHostPointer<T>> scratch_;
};

template <typename T>
class CudaLocalHostBroadcast : public LocalOp<T> {
 public:
  CudaLocalHostBroadcast(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                         CudaHostPointer<T>& source, size_t offset, size_t count);
  void runAsync() override;
  void wait() override;

 private:
  std::vector<CudaStream>& streams_;
  std::vector<CudaDevicePointer<T>> dsts_;
  CudaHostPointer<T> source_;
};

// NCCL flavours: one communicator over the distinct devices of the pointers (the reference's
// nccl::ReduceOp / BroadcastOp). Only usable when every pointer lives on a different GPU.
template <typename T>
class CudaLocalNCCLReduce : public LocalOp<T> {
 public:
  CudaLocalNCCLReduce(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                      CudaDevicePointer<T>& target, const CudaReductionFunction<T>* fn, size_t offset, size_t count);
  void runAsync() override;
  void wait() override;

 private:
  std::vector<CudaStream>& streams_;
  std::vector<CudaDevicePointer<T>> srcs_;
  CudaDevicePointer<T> target_;
  const CudaReductionFunction<T>* fn_;
  std::vector<std::shared_ptr<cuda::NcclComm>> comms_;
  int root_ = 0;
};

template <typename T>
class CudaLocalNCCLBroadcast : public LocalOp<T> {
 public:
  CudaLocalNCCLBroadcast(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                         CudaDevicePointer<T>& source, size_t offset, size_t count);
  void runAsync() override;
  void wait() override;

 private:
  std::vector<CudaStream>& streams_;
  std::vector<CudaDevicePointer<T>> dsts_;
  CudaDevicePointer<T> source_;
  std::vector<std::shared_ptr<cuda::NcclComm>> comms_;
  int root_ = 0;
};

// ---- dispatchers -------------------------------------------------------------------------------
// NCCL when the pointers sit on distinct GPUs, the library is loadable and the message is
// large enough to amortise its launch; the native kernels otherwise.
namespace cuda {
bool localOpsUseNccl(const std::vector<int>& devices, size_t bytes);
}

template <typename T>
std::unique_ptr<LocalOp<T>> cudaDeviceReduce(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                                             CudaDevicePointer<T>& targetPtr, const CudaReductionFunction<T>* fn,
                                             size_t offset, size_t count);
template <typename T>
std::unique_ptr<LocalOp<T>> cudaDeviceBroadcast(std::vector<CudaStream>& streams,
                                                std::vector<CudaDevicePointer<T>>& devicePtrs,
                                                CudaDevicePointer<T>& sourcePtr, size_t offset, size_t count);
template <typename T>
std::unique_ptr<LocalOp<T>> cudaHostReduce(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                                           CudaHostPointer<T>& targetPtr, const CudaReductionFunction<T>* fn, size_t offset,
                                           size_t count);
template <typename T>
std::unique_ptr<LocalOp<T>> cudaHostBroadcast(std::vector<CudaStream>& streams, std::vector<CudaDevicePointer<T>>& devicePtrs,
                                              CudaHostPointer<T>& sourcePtr, size_t offset, size_t count);

// Index of the pointer whose GPU is closest (PCI topology) to the transport device's NIC;
// 0 when distances are unknown. Used to pick the buffer that talks to the network.
template <typename T>
int findCudaDevicePointerClosestToDevice(std::vector<CudaDevicePointer<T>>& ptrs,
                                         const std::shared_ptr<transport::Device>& dev);

}  // namespace glb
