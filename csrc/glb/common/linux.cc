#include "glb/common/linux.h"

#include <dirent.h>
#include <limits.h>
#include <linux/ethtool.h>
#include <linux/sockios.h>
#include <net/if.h>
#include <sys/ioctl.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>

namespace glb {

const std::set<std::string>& kernelModules() {
  static std::once_flag once;
  static std::set<std::string> mods;
  std::call_once(once, [] {
    std::ifstream in("/proc/modules");
    std::string line;
    while (std::getline(in, line)) {
      auto sp = line.find(' ');
      mods.insert(line.substr(0, sp));
    }
  });
  return mods;
}

namespace {

std::string readFirstLine(const std::string& path) {
  std::ifstream in(path);
  std::string s;
  if (in) std::getline(in, s);
  return s;
}

std::vector<std::string> listDir(const std::string& path) {
  std::vector<std::string> out;
  DIR* d = opendir(path.c_str());
  if (d == nullptr) return out;
  while (struct dirent* e = readdir(d)) {
    if (e->d_name[0] == '.') continue;
    out.emplace_back(e->d_name);
  }
  closedir(d);
  return out;
}

std::vector<std::string> splitPath(const std::string& p) {
  std::vector<std::string> out;
  std::stringstream ss(p);
  std::string tok;
  while (std::getline(ss, tok, '/')) {
    if (!tok.empty()) out.push_back(tok);
  }
  return out;
}

std::string resolve(const std::string& path) {
  char buf[PATH_MAX];
  if (realpath(path.c_str(), buf) == nullptr) return "";
  return std::string(buf);
}

}  // namespace

std::vector<std::string> pciDevices(int pciClass, int mask) {
  std::vector<std::string> out;
  for (const auto& name : listDir("/sys/bus/pci/devices")) {
    auto cls = readFirstLine("/sys/bus/pci/devices/" + name + "/class");
    if (cls.empty()) continue;
    long v = std::strtol(cls.c_str(), nullptr, 16);
    if ((v & mask) == (pciClass & mask)) out.push_back(name);
  }
  return out;
}

int pciDistance(const std::string& busA, const std::string& busB) {
  auto pa = resolve("/sys/bus/pci/devices/" + busA);
  auto pb = resolve("/sys/bus/pci/devices/" + busB);
  if (pa.empty() || pb.empty()) return -1;
  auto a = splitPath(pa);
  auto b = splitPath(pb);
  size_t common = 0;
  while (common < a.size() && common < b.size() && a[common] == b[common]) common++;
  return static_cast<int>((a.size() - common) + (b.size() - common));
}

std::string interfaceToBusID(const std::string& iface) {
  auto p = resolve("/sys/class/net/" + iface + "/device");
  if (p.empty()) return "";
  auto parts = splitPath(p);
  // Walk up to the closest path component that looks like a PCI address.
  for (auto it = parts.rbegin(); it != parts.rend(); ++it) {
    unsigned dom, bus, dev, fn;
    if (std::sscanf(it->c_str(), "%x:%x:%x.%x", &dom, &bus, &dev, &fn) == 4) return *it;
  }
  return "";
}

int getInterfaceSpeedByName(const std::string& iface) {
  int sock = ::socket(AF_INET, SOCK_DGRAM, 0);
  if (sock >= 0) {
    struct ifreq ifr;
    std::memset(&ifr, 0, sizeof(ifr));
    std::strncpy(ifr.ifr_name, iface.c_str(), IFNAMSIZ - 1);
    struct ethtool_cmd ec;
    std::memset(&ec, 0, sizeof(ec));
    ec.cmd = ETHTOOL_GSET;
    ifr.ifr_data = reinterpret_cast<char*>(&ec);
    int rv = ::ioctl(sock, SIOCETHTOOL, &ifr);
    ::close(sock);
    if (rv == 0) {
      uint32_t speed = ethtool_cmd_speed(&ec);
      if (speed != 0 && speed != static_cast<uint32_t>(SPEED_UNKNOWN)) return static_cast<int>(speed);
    }
  }
  auto s = readFirstLine("/sys/class/net/" + iface + "/speed");
  if (!s.empty()) {
    int v = std::atoi(s.c_str());
    if (v > 0) return v;
  }
  return -1;
}

std::vector<std::string> listInterfaces() { return listDir("/sys/class/net"); }

}  // namespace glb
