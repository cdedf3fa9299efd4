#include "glb/cuda/stream.h"

#include <atomic>

namespace glb {
namespace cuda {

namespace {
std::mutex gDefaultMutex;
std::atomic<std::mutex*> gMutex{&gDefaultMutex};
}  // namespace

void CudaShared::setMutex(std::mutex* m) { gMutex.store(m != nullptr ? m : &gDefaultMutex); }
std::mutex& CudaShared::getMutex() { return *gMutex.load(); }

CudaStream::CudaStream(int deviceId) : deviceId_(deviceId), owner_(true) {
  DeviceGuard g(deviceId_);
  int lo = 0, hi = 0;
  GLB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  GLB_CUDA_CHECK(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, hi));
  GLB_CUDA_CHECK(cudaEventCreateWithFlags(&event_, cudaEventDisableTiming));
}

CudaStream::CudaStream(int deviceId, cudaStream_t stream) : deviceId_(deviceId), stream_(stream), owner_(false) {
  DeviceGuard g(deviceId_);
  GLB_CUDA_CHECK(cudaEventCreateWithFlags(&event_, cudaEventDisableTiming));
}

CudaStream::CudaStream(CudaStream&& o) noexcept
    : deviceId_(o.deviceId_), stream_(o.stream_), owner_(o.owner_), event_(o.event_) {
  o.stream_ = nullptr;
  o.event_ = nullptr;
  o.owner_ = false;
}

CudaStream::~CudaStream() {
  if (event_ == nullptr && stream_ == nullptr) return;
  DeviceGuard g(deviceId_);
  if (event_ != nullptr) cudaEventDestroy(event_);
  if (owner_ && stream_ != nullptr) cudaStreamDestroy(stream_);
}

void CudaStream::copyAsync(void* dst, const void* src, size_t bytes) {
  DeviceGuard g(deviceId_);
  if (bytes > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream_));
  GLB_CUDA_CHECK(cudaEventRecord(event_, stream_));
}

void CudaStream::record() {
  DeviceGuard g(deviceId_);
  GLB_CUDA_CHECK(cudaEventRecord(event_, stream_));
}

void CudaStream::wait() {
  DeviceGuard g(deviceId_);
  GLB_CUDA_CHECK(cudaEventSynchronize(event_));
}

void CudaStream::waitOn(const CudaStream& other) {
  DeviceGuard g(deviceId_);
  GLB_CUDA_CHECK(cudaStreamWaitEvent(stream_, other.event_, 0));
}

}  // namespace cuda
}  // namespace glb
