// tcp::Device: one listening socket + N epoll loops shared by every context and
// pair created from it. Inbound connections start with a 16-byte hello carrying
// the sequence number of the pair they are meant for; the device routes the
// socket to that pair (or parks it until the pair shows up).
// Parity: gloo/transport/tcp/device.{h,cc} + listener.{h,cc}.
#pragma once

#include <atomic>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "glb/transport/device.h"
#include "glb/transport/tcp/address.h"
#include "glb/transport/tcp/attr.h"
#include "glb/transport/tcp/loop.h"
#include "glb/transport/tcp/socket.h"

namespace glb {
namespace transport {
namespace tcp {

std::shared_ptr<::glb::transport::Device> CreateDevice(const struct attr&);
// Lazy devices create contexts whose pairs dial on first use instead of at rendezvous.
std::shared_ptr<::glb::transport::Device> CreateLazyDevice(const struct attr&);

struct Hello {
  static constexpr uint32_t kMagic = 0x31424c47;  // "GLB1"
  uint32_t magic = kMagic;
  uint32_t version = 1;
  sequence_number_t seq = 0;
};

class Device : public ::glb::transport::Device,
               public std::enable_shared_from_this<Device>,
               private Handler {
 public:
  using connect_callback_t = std::function<void(Socket)>;

  Device(const struct attr& attr, bool lazy);
  ~Device() override;

  std::string str() const override;
  const std::string& getPCIBusID() const override { return pciBusID_; }
  int getInterfaceSpeed() const override { return interfaceSpeed_; }
  std::shared_ptr<::glb::transport::Context> createContext(int rank, int size) override;

  bool isLazy() const { return lazy_; }
  const struct sockaddr_storage& sockaddr() const { return listenAddr_; }
  const std::string& interfaceName() const { return interfaceName_; }

  // A fresh address (listener sockaddr + unique sequence number) for a new pair.
  Address nextAddress();
  // Reserve a specific sequence number (GLB_ENABLE_RANK_AS_SEQUENCE_NUMBER mode).
  Address addressForSeq(sequence_number_t seq);

  Loop& loop(size_t i) { return *loops_[i % loops_.size()]; }
  size_t nextLoopIndex() { return loopRR_.fetch_add(1); }

  // Listener side: invoke `cb` with the socket of the inbound connection that
  // announced `seq` (immediately if it is already parked).
  void expectConnection(sequence_number_t seq, connect_callback_t cb);
  void cancelExpectation(sequence_number_t seq);

 private:
  void handleEvents(int events) override;  // accept()
  void finishHello(int fd, const Hello& hello);

  class HelloReader;
  friend class HelloReader;

  struct attr attr_;
  const bool lazy_;
  std::string interfaceName_;
  int interfaceSpeed_ = 0;
  std::string pciBusID_;

  std::vector<std::unique_ptr<Loop>> loops_;
  std::atomic<size_t> loopRR_{0};
  Socket listener_;
  struct sockaddr_storage listenAddr_ {};
  std::atomic<sequence_number_t> seq_{1ull << 32};  // keep clear of rank-as-seq values

  std::mutex mu_;
  std::unordered_map<sequence_number_t, connect_callback_t> expected_;
  std::unordered_map<sequence_number_t, Socket> parked_;
  std::map<int, std::unique_ptr<HelloReader>> readers_;
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
