// transport::ibverbs — InfiniBand / RoCE transport (reliable-connected queue pairs), the
// counterpart of gloo/transport/ibverbs/{device,pair,buffer,unbound_buffer,remote_key}.
// libibverbs is loaded at run time through a vendored ABI mirror (verbs_abi.h): the build
// needs no rdma-core headers. Protocol and design notes are in transport.cc.
//   * bound buffers: memory-region exchange + RDMA WRITE WITH IMMEDIATE;
//   * unbound buffers: eager SENDs up to 8 KB, RTS / RDMA READ / FIN rendezvous above (zero
//     copy), recv-from-any matched locally; getRemoteKey / put / get = RDMA WRITE / READ;
//   * device memory registers like host memory when nvidia_peermem is loaded (hasGPUDirect()).
// `probe()` never throws and lets callers (and the benchmark's --transport=ibverbs) fall back
// to tcp cleanly when there is no library or no HCA.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "glb/transport/device.h"

namespace glb {
namespace transport {
namespace ibverbs {

struct attr {
  std::string name;  // HCA name, e.g. "mlx5_0"; empty = first device
  int port = 1;
  int index = 0;     // GID index
};

struct Probe {
  bool libraryLoaded = false;        // libibverbs.so.1 could be dlopen'ed
  std::vector<std::string> devices;  // names returned by ibv_get_device_list
  bool peerMemoryModule = false;     // nv_peer_mem / nvidia_peermem loaded (GPUDirect RDMA)
  std::string detail;                // human-readable summary
};

// Never throws.
Probe probe();

// Names of the RDMA devices ibv_get_device_list reports (empty without libibverbs / HCA).
// Reference: gloo/transport/ibverbs/device.h getDeviceNames().
std::vector<std::string> getDeviceNames();

// Opens the HCA (first device when `name` is empty). Throws InvalidOperationException naming
// the missing piece when libibverbs or an RDMA device is absent.
std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr&);

}  // namespace ibverbs
}  // namespace transport
}  // namespace glb
