// CudaStream: a high-priority non-blocking stream (owned or borrowed) plus one
// timing-disabled event recorded after every async copy; wait() blocks the host
// on that event. CudaDevicePointer / CudaHostPointer: typed (ptr, count, device)
// handles, owning or not. Parity: gloo/cuda.h:56-272, cuda.cu:40-246.
#pragma once

#include <cuda_runtime.h>

#include <memory>
#include <mutex>

#include "glb/cuda/cuda_util.h"

namespace glb {
namespace cuda {

// Process-wide mutex serialising cudaMalloc / cudaFree / cudaMallocHost against
// NCCL launches (the deadlock the reference documents in docs/cuda.md:40-59).
// Applications that already hold such a lock can install their own.
class CudaShared {
 public:
  static void setMutex(std::mutex* m);
  static std::mutex& getMutex();
};

class CudaStream {
 public:
  // Creates a new high-priority non-blocking stream on `deviceId` ...
  explicit CudaStream(int deviceId);
  // ... or wraps an existing one (not destroyed).
  CudaStream(int deviceId, cudaStream_t stream);
  CudaStream(CudaStream&& other) noexcept;
  CudaStream(const CudaStream&) = delete;
  ~CudaStream();

  cudaStream_t operator*() const { return stream_; }
  cudaStream_t get() const { return stream_; }
  int getDeviceID() const { return deviceId_; }
  cudaEvent_t getEvent() const { return event_; }

  // dst/src may each be device or pinned host memory.
  void copyAsync(void* dst, const void* src, size_t bytes);
  void record();
  void wait();                      // host blocks until the last recorded point
  void waitOn(const CudaStream& other);  // this stream waits for other's last recorded point

 private:
  int deviceId_;
  cudaStream_t stream_ = nullptr;
  bool owner_ = false;
  cudaEvent_t event_ = nullptr;
};

template <typename T>
class CudaDevicePointer {
 public:
  static CudaDevicePointer<T> alloc(size_t count) {
    T* p = nullptr;
    {
      std::lock_guard<std::mutex> g(CudaShared::getMutex());
      GLB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
    }
    CudaDevicePointer<T> out(p, count, true);
    return out;
  }
  static CudaDevicePointer<T> create(T* ptr, size_t count) { return CudaDevicePointer<T>(ptr, count, false); }
  static CudaDevicePointer<T> create(const CudaDevicePointer<T>& o) { return create(o.device_, o.count_); }

  CudaDevicePointer() = default;
  CudaDevicePointer(CudaDevicePointer&& o) noexcept { *this = std::move(o); }
  CudaDevicePointer& operator=(CudaDevicePointer&& o) noexcept {
    release();
    device_ = o.device_;
    count_ = o.count_;
    owner_ = o.owner_;
    deviceId_ = o.deviceId_;
    o.device_ = nullptr;
    o.owner_ = false;
    o.count_ = 0;
    return *this;
  }
  CudaDevicePointer(const CudaDevicePointer&) = delete;
  ~CudaDevicePointer() { release(); }

  T* operator*() const { return device_; }
  T& operator[](size_t i) const { return device_[i]; }
  size_t getCount() const { return count_; }
  int getDeviceID() const { return deviceId_; }
  CudaDevicePointer<T> range(size_t offset, size_t count) const {
    GLB_ENFORCE_LE(offset + count, count_);
    return CudaDevicePointer<T>(device_ + offset, count, false);
  }

 private:
  CudaDevicePointer(T* ptr, size_t count, bool owner) : device_(ptr), count_(count), owner_(owner) {
    deviceId_ = ptr != nullptr ? deviceForPointer(ptr) : -1;
  }
  void release() {
    if (owner_ && device_ != nullptr) {
      std::lock_guard<std::mutex> g(CudaShared::getMutex());
      cudaFree(device_);
    }
    device_ = nullptr;
  }
  T* device_ = nullptr;
  size_t count_ = 0;
  bool owner_ = false;
  int deviceId_ = -1;
};

template <typename T>
class CudaHostPointer {
 public:
  static CudaHostPointer<T> alloc(size_t count) {
    T* p = nullptr;
    {
      std::lock_guard<std::mutex> g(CudaShared::getMutex());
      GLB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&p), std::max<size_t>(1, count) * sizeof(T)));
    }
    return CudaHostPointer<T>(p, count, true);
  }
  static CudaHostPointer<T> create(T* ptr, size_t count) { return CudaHostPointer<T>(ptr, count, false); }

  CudaHostPointer() = default;
  CudaHostPointer(CudaHostPointer&& o) noexcept { *this = std::move(o); }
  CudaHostPointer& operator=(CudaHostPointer&& o) noexcept {
    release();
    host_ = o.host_;
    count_ = o.count_;
    owner_ = o.owner_;
    o.host_ = nullptr;
    o.owner_ = false;
    o.count_ = 0;
    return *this;
  }
  CudaHostPointer(const CudaHostPointer&) = delete;
  ~CudaHostPointer() { release(); }

  T* operator*() const { return host_; }
  T& operator[](size_t i) const { return host_[i]; }
  size_t getCount() const { return count_; }
  CudaHostPointer<T> range(size_t offset, size_t count) const {
    GLB_ENFORCE_LE(offset + count, count_);
    return CudaHostPointer<T>(host_ + offset, count, false);
  }

 private:
  CudaHostPointer(T* ptr, size_t count, bool owner) : host_(ptr), count_(count), owner_(owner) {}
  void release() {
    if (owner_ && host_ != nullptr) {
      std::lock_guard<std::mutex> g(CudaShared::getMutex());
      cudaFreeHost(host_);
    }
    host_ = nullptr;
  }
  T* host_ = nullptr;
  size_t count_ = 0;
  bool owner_ = false;
};

}  // namespace cuda
}  // namespace glb
