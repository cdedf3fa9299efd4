// Linux sysfs / procfs probes: kernel modules, PCI device enumeration, PCI
// distance, network interface -> PCI bus id and link speed.
// Parity: gloo/common/linux.{h,cc}, linux_devices.h. The reference's topology
// knowledge ends at PCI; NVLink discovery lives in cuda/topology.{h,cc}.
#pragma once

#include <set>
#include <string>
#include <vector>

namespace glb {

constexpr int kPCIClass3D = 0x030200;       // 3D controller (GPUs in SXM boxes)
constexpr int kPCIClassVGA = 0x030000;      // display controller
constexpr int kPCIClassNetwork = 0x020000;  // network controller (top byte 0x02)
constexpr int kPCIClassBridgeNVSwitch = 0x068000;

const std::set<std::string>& kernelModules();

// Bus ids ("0000:1b:00.0") of PCI devices whose class matches `pciClass`
// under the mask (0xffff00 compares class+subclass, 0xff0000 class only).
std::vector<std::string> pciDevices(int pciClass, int mask = 0xffff00);

// Number of PCI hops separating two devices: 0 when identical, otherwise the
// count of non-shared path components of their sysfs device paths. -1 when
// either device is unknown.
int pciDistance(const std::string& busA, const std::string& busB);

// PCI bus id backing a network interface ("" for virtual interfaces like lo).
std::string interfaceToBusID(const std::string& iface);

// Link speed in Mb/s from ethtool (falls back to /sys/class/net/<if>/speed); -1 if unknown.
int getInterfaceSpeedByName(const std::string& iface);

// All interface names on the host.
std::vector<std::string> listInterfaces();

}  // namespace glb
