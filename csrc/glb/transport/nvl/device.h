// transport::nvl — the NVLink / NVSwitch peer-memory data plane behind the transport API.
//
// In the reference, GPU memory reaches a transport only through ibverbs + GPUDirect RDMA:
// Device::hasGPUDirect() (ibverbs/device.cc:187-189), a device pointer as UnboundBuffer
// (ibverbs/unbound_buffer.cc:215-235) and one-sided put / get through a RemoteKey
// (ibverbs/pair.cc:408-530). This device offers the same surface for GPUs of one NVLink
// domain:
//
//   * host pointers behave exactly as on the wrapped control-plane device (tcp / tls): the
//     nvl context forwards pairs, rendezvous and host unbound buffers to it;
//   * a DEVICE pointer handed to createUnboundBuffer() becomes an nvl::UnboundBuffer:
//       send / recv      kernels that stream the payload through the receiver's mailbox ring
//                        over NVLink (cuda/p2p_kernels.cu); waitSend / waitRecv complete them.
//                        Matching is by posting order per (source, destination) pair.
//       getRemoteKey     exports the allocation as a CUDA IPC handle (+ offset); the key is a
//                        string and can travel through any collective or store;
//       put / get        the initiator maps the target lazily (cached per handle) and runs a
//                        copy kernel against the peer pointer: one-sided, the owner does
//                        nothing. Completion: waitSend (put) / waitRecv (get).
//   * Device::hasGPUDirect() == true.
//
// The context creates the CUDA PeerContext (collectively) when it is attached to its
// glb::Context at the end of connectFullMesh, so every rank must use an nvl device.
#pragma once

#include <memory>
#include <string>

#include "glb/transport/device.h"

namespace glb {
namespace transport {
namespace nvl {

struct attr {
  std::shared_ptr<::glb::transport::Device> control;  // host transport (tcp / tls device)
  int cudaDevice = -1;                                // -1: the calling thread's current device
};

std::shared_ptr<::glb::transport::Device> CreateDevice(const attr&);

}  // namespace nvl
}  // namespace transport
}  // namespace glb
