// A transport device: factory for per-communicator transport contexts; may own
// I/O threads shared by all of them. Parity: gloo/transport/device.h:34-54.
#pragma once

#include <memory>
#include <string>

namespace glb {
namespace transport {

class Context;

class Device {
 public:
  virtual ~Device() = default;

  virtual std::string str() const = 0;
  virtual const std::string& getPCIBusID() const = 0;
  virtual int getInterfaceSpeed() const { return 0; }
  // True when device memory can be handed to this transport without staging
  // (reference: GPUDirect RDMA; here also the NVLink peer-memory device).
  virtual bool hasGPUDirect() const { return false; }

  virtual std::shared_ptr<Context> createContext(int rank, int size) = 0;
};

}  // namespace transport
}  // namespace glb
