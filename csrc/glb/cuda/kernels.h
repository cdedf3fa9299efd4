// Host-callable launchers for the device kernels (implemented in the .cu files).
#pragma once

#include <cuda_runtime.h>

#include "glb/cuda/comm_types.h"
#include "glb/types.h"

namespace glb {
namespace cuda {

// Force-load all kernels of the library on the current device (idempotent, cheap).
void preloadAllreduceKernels();
void preloadCollectiveKernels();
void preloadScheduleKernels();
void preloadLocalKernels();

// allreduce_kernels.cu
void launchBarrier(const CommArgs& a, cudaStream_t stream);
void launchOneShotAllreduce(const CommArgs& a, const void* in, void* out, size_t count, DataType dt, ReduceOp op,
                            const PeerPtrs& stage, size_t halfBytes, int blocks, cudaStream_t stream);
void launchTwoShotAllreduce(const CommArgs& a, const PeerPtrs& bufs, size_t count, DataType dt, ReduceOp op,
                            bool vectorOk, int blocks, cudaStream_t stream);
bool nvlsSupports(DataType dt, ReduceOp op);
void setOneShotPush(bool on);  // tuning / A-B testing of the push flavour of one-shot
bool oneShotPushEnabled();
void launchNvlsAllreduce(const CommArgs& a, void* mcPtr, const PeerPtrs& bufs, size_t count, DataType dt, int blocks,
                         cudaStream_t stream);

// collective_kernels.cu
void launchBroadcast(const CommArgs& a, const PeerPtrs& bufs, void* mc, size_t bytes, int root, int mode, bool vec,
                     int blocks, cudaStream_t stream);
void launchGatherPush(const CommArgs& a, const void* in, const PeerPtrs& outs, void* mcOut, const size_t* offs,
                      const size_t* lens, int onlyDst, bool vec, int blocks, cudaStream_t stream);
void launchAlltoallPush(const CommArgs& a, const void* in, const PeerPtrs& outs, const size_t* sendOff,
                        const size_t* sendLen, const size_t* dstOff, const size_t* recvOffTable, int onlySrc,
                        bool vec, int blocks, cudaStream_t stream);
void launchReducePull(const CommArgs& a, const PeerPtrs& ins, void* mcIn, void* out, const size_t* elemOff,
                      const size_t* elemLen, DataType dt, ReduceOp op, bool vec, bool useMc, int blocks,
                      cudaStream_t stream);

// reduce_kernels.cu — local element-wise ops: dst = dst (op) src, and
// multi-source reduce / broadcast between buffers visible to one device.
void launchLocalReduce(void* dst, const void* src, size_t count, DataType dt, ReduceOp op, cudaStream_t stream);
void launchLocalReduceMany(void* dst, const void* const* srcs, int nsrc, size_t count, DataType dt, ReduceOp op,
                           cudaStream_t stream);
void launchLocalBroadcast(void* const* dsts, int ndst, const void* src, size_t bytes, cudaStream_t stream);
void launchFill(void* dst, size_t count, DataType dt, double start, double stride, cudaStream_t stream);
void launchSpin(long long cycles, cudaStream_t stream);

}  // namespace cuda
}  // namespace glb
