#include "glb/cuda/cuda_util.h"

#include <cstring>
#include <mutex>

namespace glb {
namespace cuda {

int deviceCount() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear sticky "no device" state
    return 0;
  }
  return n;
}

int currentDevice() {
  int d = 0;
  GLB_CUDA_CHECK(cudaGetDevice(&d));
  return d;
}

int deviceForPointer(const void* ptr) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return attr.device;
  return -1;
}

std::string devicePCIBusId(int device) {
  char buf[32] = {0};
  GLB_CUDA_CHECK(cudaDeviceGetPCIBusId(buf, sizeof(buf), device));
  for (char* c = buf; *c; ++c) *c = static_cast<char>(std::tolower(*c));
  return std::string(buf);
}

std::string deviceUUID(int device) {
  cudaDeviceProp prop;
  GLB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  return std::string(reinterpret_cast<const char*>(prop.uuid.bytes), 16);
}

std::string driverErrorString(CUresult r) {
  try {
    const char* s = nullptr;
    if (driver().cuGetErrorString(r, &s) == CUDA_SUCCESS && s != nullptr) return s;
  } catch (...) {
  }
  return "unknown";
}

namespace {
template <typename F>
bool load(F& fn, const char* name, bool required) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    if (required) GLB_THROW(Exception, "CUDA driver entry point not available: ", name);
    fn = nullptr;
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}
}  // namespace

const DriverApi& driver() {
  static DriverApi api;
  static std::once_flag once;
  static std::string error;
  std::call_once(once, [] {
    try {
      GLB_CUDA_CHECK(cudaFree(nullptr));  // make sure the runtime (and driver) are initialised
      load(api.cuGetErrorString, "cuGetErrorString", true);
      load(api.cuDeviceGet, "cuDeviceGet", true);
      load(api.cuDeviceGetAttribute, "cuDeviceGetAttribute", true);
      load(api.cuMemGetAllocationGranularity, "cuMemGetAllocationGranularity", true);
      load(api.cuMemCreate, "cuMemCreate", true);
      load(api.cuMemRelease, "cuMemRelease", true);
      load(api.cuMemAddressReserve, "cuMemAddressReserve", true);
      load(api.cuMemAddressFree, "cuMemAddressFree", true);
      load(api.cuMemMap, "cuMemMap", true);
      load(api.cuMemUnmap, "cuMemUnmap", true);
      load(api.cuMemSetAccess, "cuMemSetAccess", true);
      load(api.cuMemExportToShareableHandle, "cuMemExportToShareableHandle", true);
      load(api.cuMemImportFromShareableHandle, "cuMemImportFromShareableHandle", true);
      load(api.cuMemGetAddressRange, "cuMemGetAddressRange", true);
      bool mc = load(api.cuMulticastCreate, "cuMulticastCreate", false);
      mc = load(api.cuMulticastAddDevice, "cuMulticastAddDevice", false) && mc;
      mc = load(api.cuMulticastBindMem, "cuMulticastBindMem", false) && mc;
      mc = load(api.cuMulticastUnbind, "cuMulticastUnbind", false) && mc;
      mc = load(api.cuMulticastGetGranularity, "cuMulticastGetGranularity", false) && mc;
      api.haveMulticast = mc;
    } catch (const std::exception& e) {
      error = e.what();
    }
  });
  if (!error.empty()) GLB_THROW(Exception, "CUDA driver unavailable: ", error);
  return api;
}

}  // namespace cuda
}  // namespace glb
