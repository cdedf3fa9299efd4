// CUDA helper macros, device guards and a lazily resolved driver-API table.
// The driver entry points (cuMem*, cuMulticast*) are fetched through
// cudaGetDriverEntryPoint so the library has no link-time dependency on
// libcuda.so and still loads on hosts without a GPU.
// Parity: gloo/cuda_private.{h,cu} (CUDA_CHECK, CudaDeviceGuard/Scope, getGPUIDForPointer).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "glb/common/logging.h"

#define GLB_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t glb_err_ = (expr);                                                            \
    if (glb_err_ != cudaSuccess) {                                                            \
      GLB_THROW(::glb::Exception, "CUDA error: ", cudaGetErrorName(glb_err_), " (",           \
                cudaGetErrorString(glb_err_), ") at " #expr);                                 \
    }                                                                                         \
  } while (0)

#define GLB_CU_CHECK(expr)                                                                    \
  do {                                                                                        \
    CUresult glb_res_ = (expr);                                                               \
    if (glb_res_ != CUDA_SUCCESS) {                                                           \
      GLB_THROW(::glb::Exception, "CUDA driver error ", static_cast<int>(glb_res_), " (",     \
                ::glb::cuda::driverErrorString(glb_res_), ") at " #expr);                     \
    }                                                                                         \
  } while (0)

namespace glb {
namespace cuda {

// Number of visible devices; 0 when there is no driver / GPU (never throws).
int deviceCount();
int currentDevice();
int deviceForPointer(const void* ptr);  // -1 for non-device pointers
std::string devicePCIBusId(int device);
std::string deviceUUID(int device);  // 16 raw bytes
std::string driverErrorString(CUresult r);

class DeviceGuard {  // restores the current device on scope exit
 public:
  DeviceGuard() { cudaGetDevice(&prev_); }
  explicit DeviceGuard(int device) {
    cudaGetDevice(&prev_);
    if (device >= 0 && device != prev_) GLB_CUDA_CHECK(cudaSetDevice(device));
  }
  ~DeviceGuard() { cudaSetDevice(prev_); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;

 private:
  int prev_ = 0;
};
using DeviceScope = DeviceGuard;

// Driver API entry points used by the VMM / multicast allocator.
struct DriverApi {
  CUresult (*cuGetErrorString)(CUresult, const char**);
  CUresult (*cuDeviceGet)(CUdevice*, int);
  CUresult (*cuDeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*cuMemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*cuMemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*cuMemRelease)(CUmemGenericAllocationHandle);
  CUresult (*cuMemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*cuMemAddressFree)(CUdeviceptr, size_t);
  CUresult (*cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*cuMemUnmap)(CUdeviceptr, size_t);
  CUresult (*cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*cuMemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*cuMemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*cuMemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr);
  CUresult (*cuMulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*cuMulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*cuMulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*cuMulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*cuMulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  bool haveMulticast = false;
};

// Throws if the driver is unavailable.
const DriverApi& driver();

}  // namespace cuda
}  // namespace glb
