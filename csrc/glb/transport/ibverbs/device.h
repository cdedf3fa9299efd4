// transport::ibverbs — availability probe and API surface of the reference's InfiniBand
// transport (gloo/transport/ibverbs/device.h:24-60).
//
// What the reference used it for — zero-copy one-sided transfers and handing GPU memory to
// the transport without staging — is served here by two other pieces that exist and are
// tested: `UnboundBuffer::getRemoteKey/put/get` on the tcp transport (software one-sided,
// single-copy between ranks of one host) and `cuda::PeerContext` (NVLink / NVSwitch peer
// memory, NVLS multicast). A verbs data path is NOT implemented: the image ships neither
// the rdma-core headers nor an HCA to run it against, and a B200 HGX box reaches its peers
// over NVSwitch, not over a NIC. `CreateDevice` therefore reports precisely what is missing
// instead of pretending; `probe()` lets callers (and the benchmark's --transport=ibverbs)
// fall back to tcp cleanly.
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

// Throws InvalidOperationException naming the missing piece (library, device, or the
// verbs data path of this build).
std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr&);

}  // namespace ibverbs
}  // namespace transport
}  // namespace glb
