#include "glb/transport/ibverbs/device.h"

#include <dlfcn.h>

#include "glb/common/error.h"
#include "glb/common/linux.h"
#include "glb/common/logging.h"
#include "glb/common/string.h"

namespace glb {
namespace transport {
namespace ibverbs {

namespace {
// The two entry points needed to enumerate devices are plain exported functions with a
// stable ABI (device structs are only passed back opaquely, names come from an accessor).
struct ibv_device;
using GetDeviceList = ibv_device** (*)(int*);
using FreeDeviceList = void (*)(ibv_device**);
using GetDeviceName = const char* (*)(ibv_device*);
}  // namespace

Probe probe() {
  Probe p;
  const auto& mods = kernelModules();
  p.peerMemoryModule = mods.count("nv_peer_mem") > 0 || mods.count("nvidia_peermem") > 0;
  void* lib = ::dlopen("libibverbs.so.1", RTLD_NOW | RTLD_LOCAL);
  if (lib == nullptr) {
    p.detail = "libibverbs.so.1 is not installed";
    return p;
  }
  p.libraryLoaded = true;
  auto getList = reinterpret_cast<GetDeviceList>(::dlsym(lib, "ibv_get_device_list"));
  auto freeList = reinterpret_cast<FreeDeviceList>(::dlsym(lib, "ibv_free_device_list"));
  auto getName = reinterpret_cast<GetDeviceName>(::dlsym(lib, "ibv_get_device_name"));
  if (getList != nullptr && freeList != nullptr && getName != nullptr) {
    int n = 0;
    ibv_device** list = getList(&n);
    if (list != nullptr) {
      for (int i = 0; i < n; i++) {
        const char* name = getName(list[i]);
        if (name != nullptr) p.devices.emplace_back(name);
      }
      freeList(list);
    }
  }
  ::dlclose(lib);
  p.detail = p.devices.empty() ? "libibverbs is present but no RDMA device was found"
                               : strcat_all(p.devices.size(), " RDMA device(s) found");
  return p;
}

std::vector<std::string> getDeviceNames() { return probe().devices; }

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr& a) {
  Probe p = probe();
  if (!p.libraryLoaded || p.devices.empty()) {
    GLB_THROW_INVALID_OPERATION_EXCEPTION("ibverbs transport unavailable: ", p.detail,
                                ". Use transport::tcp (same-host ranks get a single-copy path) or, for "
                                "CUDA buffers, cuda::PeerContext over NVLink.");
  }
  GLB_THROW_INVALID_OPERATION_EXCEPTION("ibverbs transport: device '", a.name.empty() ? p.devices.front() : a.name,
                              "' exists but this build has no verbs data path (see transport/ibverbs/device.h); "
                              "use transport::tcp or cuda::PeerContext.");
}

}  // namespace ibverbs
}  // namespace transport
}  // namespace glb
