// transport::ibverbs — InfiniBand / RoCE transport over reliable-connected queue pairs.
//
// Same capabilities as the reference's ibverbs transport (RC QP setup pair.cc:69-191, memory
// region exchange :222-281, RDMA writes :328-384, one-sided put/get :408-530, completion
// thread device.cc:122-235, GPUDirect registration of device memory buffer.cc:24-59), with a
// wire protocol of its own that mirrors this library's TCP transport:
//
//   * every pair pre-posts a ring of receive slots; small control messages travel as SENDs
//     with the 64-byte header INLINE (no buffer to manage, usable from the completion thread);
//   * unbound send <= 8 KB of host memory: EAGER (header + payload through a registered
//     bounce slot), matched against posted receives or parked in the unexpected queue;
//   * larger / device-memory unbound sends: RTS {address, rkey, length}; the receiver issues
//     an RDMA READ straight into the destination once a receive matches and answers FIN —
//     zero copies, and recv-from-any is a purely receiver-local match;
//   * bound buffers (old-style algorithms): the receive side publishes its memory region
//     (MR message), the send side does RDMA WRITE WITH IMMEDIATE (imm = slot);
//   * RemoteKey put / get: RDMA WRITE / READ against {address, rkey} of the key.
//
// libibverbs is loaded at run time (verbs_abi.h); a device pointer is registered like any
// other buffer, which works when nvidia_peermem is loaded (`hasGPUDirect()`).
#include <arpa/inet.h>
#include <dlfcn.h>
#include <poll.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>

#include "glb/common/error.h"
#include "glb/common/linux.h"
#include "glb/common/logging.h"
#include "glb/common/string.h"
#include "glb/common/utils.h"
#include "glb/cuda/cuda_util.h"
#include "glb/transport/context.h"
#include "glb/transport/ibverbs/device.h"
#include "glb/transport/ibverbs/verbs_abi.h"

namespace glb {
namespace transport {
namespace ibverbs {

// ---- library loading ------------------------------------------------------------------------

const VerbsApi* verbs(std::string* why) {
  static VerbsApi api;
  static std::string error;
  static const bool ok = [] {
    const char* forced = std::getenv("GLB_IBVERBS_LIB");
    void* lib = ::dlopen(forced != nullptr ? forced : "libibverbs.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (lib == nullptr) {
      error = strcat_all(forced != nullptr ? forced : "libibverbs.so.1", " is not installed");
      return false;
    }
    bool all = true;
#define GLB_IBV(field, sym)                                                     \
  api.field = reinterpret_cast<decltype(api.field)>(::dlsym(lib, sym));         \
  if (api.field == nullptr) {                                                   \
    all = false;                                                                \
    error = "libibverbs lacks " sym;                                            \
  }
    GLB_IBV(get_device_list, "ibv_get_device_list")
    GLB_IBV(free_device_list, "ibv_free_device_list")
    GLB_IBV(get_device_name, "ibv_get_device_name")
    GLB_IBV(open_device, "ibv_open_device")
    GLB_IBV(close_device, "ibv_close_device")
    GLB_IBV(alloc_pd, "ibv_alloc_pd")
    GLB_IBV(dealloc_pd, "ibv_dealloc_pd")
    GLB_IBV(reg_mr, "ibv_reg_mr")
    GLB_IBV(dereg_mr, "ibv_dereg_mr")
    GLB_IBV(create_comp_channel, "ibv_create_comp_channel")
    GLB_IBV(destroy_comp_channel, "ibv_destroy_comp_channel")
    GLB_IBV(create_cq, "ibv_create_cq")
    GLB_IBV(destroy_cq, "ibv_destroy_cq")
    GLB_IBV(get_cq_event, "ibv_get_cq_event")
    GLB_IBV(ack_cq_events, "ibv_ack_cq_events")
    GLB_IBV(create_qp, "ibv_create_qp")
    GLB_IBV(destroy_qp, "ibv_destroy_qp")
    GLB_IBV(modify_qp, "ibv_modify_qp")
    GLB_IBV(query_port, "ibv_query_port")
    GLB_IBV(query_gid, "ibv_query_gid")
#undef GLB_IBV
    return all;
  }();
  if (!ok) {
    if (why != nullptr) *why = error;
    return nullptr;
  }
  return &api;
}

namespace {

constexpr size_t kEagerMax = 8 * 1024;
constexpr int kRecvSlots = 128;   // pre-posted receive slots per pair
constexpr int kSendSlots = 32;    // bounce slots for eager payloads per pair
constexpr int kSendDepth = 256;   // send queue depth
constexpr uint32_t kInline = 64;

enum MsgType : uint32_t { MSG_EAGER = 1, MSG_RTS = 2, MSG_FIN = 3, MSG_MR = 4 };

struct Header {  // travels inline (64 bytes)
  uint32_t type;
  uint32_t id;       // RTS / FIN: the sender's operation id
  uint64_t slot;
  uint64_t nbytes;
  uint64_t raddr;    // RTS: source address, MR: region address
  uint32_t rkey;
  uint32_t pad;
  uint64_t reserved[3];
};
static_assert(sizeof(Header) == 64, "header must fit the inline budget");

class Dev;
class Ctx;
class PairImpl;
class UBuf;

struct Mr {
  const VerbsApi* api = nullptr;
  ibv_mr* mr = nullptr;
  ~Mr() {
    if (mr != nullptr) api->dereg_mr(mr);
  }
};

// What a send-side completion refers to.
struct Op {
  enum Kind { CONTROL, EAGER_SLOT, BOUND_WRITE, READ_FOR_RECV, PUT, GET } kind = CONTROL;
  int sendSlot = -1;        // EAGER_SLOT
  class BoundBuf* bound = nullptr;
  UBuf* ubuf = nullptr;     // READ_FOR_RECV / PUT / GET
  uint32_t id = 0;          // READ_FOR_RECV: the peer's op id to FIN
};

class Addr : public Address {
 public:
  struct Blob {
    uint16_t lid;
    uint16_t pad;
    uint32_t qpn;
    uint32_t psn;
    uint32_t mtu;
    uint8_t gid[16];
  } b{};
  std::string str() const override { return strcat_all("ibv lid=", b.lid, " qpn=", b.qpn, " psn=", b.psn); }
  std::vector<char> bytes() const override {
    std::vector<char> v(sizeof(b));
    std::memcpy(v.data(), &b, sizeof(b));
    return v;
  }
};

class Key : public RemoteKey {
 public:
  Key(int rank, uint64_t addr, uint32_t rkey, size_t size) : RemoteKey(rank, size), addr(addr), rkey(rkey) {}
  std::string serialize() const override { return strcat_all("ibv:", rank, ":", addr, ":", rkey, ":", size); }
  uint64_t addr;
  uint32_t rkey;
};

// ---- device: HCA, protection domain, completion queue + thread ---------------------------------

class Dev : public Device, public std::enable_shared_from_this<Dev> {
 public:
  Dev(const attr& a, const VerbsApi* api) : api_(api), port_(a.port), gidIndex_(a.index) {
    int n = 0;
    ibv_device** list = api_->get_device_list(&n);
    GLB_ENFORCE(list != nullptr && n > 0, "no RDMA device found");
    ibv_device* chosen = nullptr;
    for (int i = 0; i < n; i++) {
      const char* nm = api_->get_device_name(list[i]);
      if (a.name.empty() || (nm != nullptr && a.name == nm)) {
        chosen = list[i];
        name_ = nm != nullptr ? nm : "?";
        break;
      }
    }
    if (chosen != nullptr) ctx_ = api_->open_device(chosen);
    api_->free_device_list(list);
    GLB_ENFORCE(ctx_ != nullptr, "cannot open RDMA device '", a.name, "'");
    pd_ = api_->alloc_pd(ctx_);
    GLB_ENFORCE(pd_ != nullptr, "ibv_alloc_pd failed");
    channel_ = api_->create_comp_channel(ctx_);
    GLB_ENFORCE(channel_ != nullptr, "ibv_create_comp_channel failed");
    cq_ = api_->create_cq(ctx_, 8192, nullptr, channel_, 0);
    GLB_ENFORCE(cq_ != nullptr, "ibv_create_cq failed");
    ibv_port_attr pa;
    std::memset(&pa, 0, sizeof(pa));
    GLB_ENFORCE_EQ(api_->query_port(ctx_, static_cast<uint8_t>(port_), &pa), 0, "ibv_query_port failed");
    lid_ = pa.lid;
    mtu_ = pa.active_mtu != 0 ? pa.active_mtu : IBV_MTU_1024;
    std::memset(&gid_, 0, sizeof(gid_));
    api_->query_gid(ctx_, static_cast<uint8_t>(port_), gidIndex_, &gid_);
    GLB_ENFORCE_EQ(ctx_->ops.req_notify_cq(cq_, 0), 0, "ibv_req_notify_cq failed");
    const auto& mods = kernelModules();
    gpuDirect_ = mods.count("nv_peer_mem") > 0 || mods.count("nvidia_peermem") > 0;
    thread_ = std::thread([this] { loop(); });
  }
  ~Dev() override {
    done_ = true;
    if (thread_.joinable()) thread_.join();
    if (cq_ != nullptr) api_->destroy_cq(cq_);
    if (channel_ != nullptr) api_->destroy_comp_channel(channel_);
    if (pd_ != nullptr) api_->dealloc_pd(pd_);
    if (ctx_ != nullptr) api_->close_device(ctx_);
  }

  std::string str() const override { return strcat_all("ibverbs(", name_, ":", port_, " gid ", gidIndex_, ")"); }
  const std::string& getPCIBusID() const override { return pci_; }
  bool hasGPUDirect() const override { return gpuDirect_; }
  std::shared_ptr<Context> createContext(int rank, int size) override;

  const VerbsApi* api() const { return api_; }
  ibv_context* ctx() const { return ctx_; }
  ibv_pd* pd() const { return pd_; }
  ibv_cq* cq() const { return cq_; }
  int port() const { return port_; }
  int gidIndex() const { return gidIndex_; }
  uint16_t lid() const { return lid_; }
  ibv_mtu mtu() const { return mtu_; }
  const ibv_gid& gid() const { return gid_; }

  std::shared_ptr<Mr> reg(void* ptr, size_t size, int access) {
    auto m = std::make_shared<Mr>();
    m->api = api_;
    m->mr = api_->reg_mr(pd_, ptr, std::max<size_t>(size, 1), access);
    if (m->mr == nullptr) {
      GLB_THROW_IO_EXCEPTION("ibv_reg_mr(", size, " bytes) failed: ", std::strerror(errno),
                             gpuDirect_ ? "" : " (device memory needs the nvidia_peermem module)");
    }
    return m;
  }

  void addPair(uint32_t qpn, PairImpl* p) {
    std::lock_guard<std::mutex> g(mu_);
    pairs_[qpn] = p;
  }
  void removePair(uint32_t qpn) {
    std::unique_lock<std::mutex> g(mu_);
    pairs_.erase(qpn);
    // the completion thread may be inside the pair right now
    while (busyQpn_ == qpn && std::this_thread::get_id() != thread_.get_id()) {
      g.unlock();
      std::this_thread::yield();
      g.lock();
    }
  }

 private:
  void loop();
  void drain();

  const VerbsApi* api_;
  int port_, gidIndex_;
  std::string name_, pci_;
  ibv_context* ctx_ = nullptr;
  ibv_pd* pd_ = nullptr;
  ibv_comp_channel* channel_ = nullptr;
  ibv_cq* cq_ = nullptr;
  uint16_t lid_ = 0;
  ibv_mtu mtu_ = IBV_MTU_1024;
  ibv_gid gid_;
  bool gpuDirect_ = false;
  std::atomic<bool> done_{false};
  std::thread thread_;
  std::mutex mu_;
  std::map<uint32_t, PairImpl*> pairs_;
  uint32_t busyQpn_ = 0;
};

// ---- bound buffers -----------------------------------------------------------------------------

class BoundBuf : public Buffer {
 public:
  BoundBuf(PairImpl* pair, int slot, void* ptr, size_t size, bool isRecv, std::shared_ptr<Mr> mr)
      : Buffer(slot, ptr, size), pair_(pair), isRecv_(isRecv), mr_(std::move(mr)) {}
  ~BoundBuf() override;
  void send(size_t offset, size_t length, size_t roffset) override;
  void waitRecv() override { wait(recvDone_, "recv"); }
  void waitSend() override { wait(sendDone_, "send"); }
  void onRecv() { bump(recvDone_); }
  void onSend() { bump(sendDone_); }
  void fail(const std::string& why) {
    std::lock_guard<std::mutex> g(mu_);
    error_ = why;
    cv_.notify_all();
  }
  uint32_t lkey() const { return mr_->mr->lkey; }
  uint32_t rkey() const { return mr_->mr->rkey; }

 private:
  void bump(int& counter) {
    std::lock_guard<std::mutex> g(mu_);
    counter++;
    cv_.notify_all();
  }
  void wait(int& counter, const char* what);
  PairImpl* pair_;
  bool isRecv_;
  std::shared_ptr<Mr> mr_;
  std::mutex mu_;
  std::condition_variable cv_;
  int recvDone_ = 0, sendDone_ = 0;
  std::string error_;
};

// ---- unbound buffer ----------------------------------------------------------------------------

class UBuf : public UnboundBuffer {
 public:
  UBuf(Ctx* ctx, void* ptr, size_t size);
  ~UBuf() override;
  bool waitRecv(int* rank, std::chrono::milliseconds timeout) override { return wait(recvs_, abortRecv_, rank, timeout, "recv"); }
  bool waitSend(int* rank, std::chrono::milliseconds timeout) override { return wait(sends_, abortSend_, rank, timeout, "send"); }
  void abortWaitRecv() override {
    std::lock_guard<std::mutex> g(mu_);
    abortRecv_ = true;
    cv_.notify_all();
  }
  void abortWaitSend() override {
    std::lock_guard<std::mutex> g(mu_);
    abortSend_ = true;
    cv_.notify_all();
  }
  void send(int dstRank, uint64_t slot, size_t offset, size_t nbytes) override;
  void recv(int srcRank, uint64_t slot, size_t offset, size_t nbytes) override { recv(std::vector<int>{srcRank}, slot, offset, nbytes); }
  void recv(std::vector<int> srcRanks, uint64_t slot, size_t offset, size_t nbytes) override;
  std::unique_ptr<RemoteKey> getRemoteKey() const override;
  void put(const RemoteKey& key, uint64_t slot, size_t offset, size_t roffset, size_t nbytes) override;
  void get(const RemoteKey& key, uint64_t slot, size_t offset, size_t roffset, size_t nbytes) override;

  void onSend(int rank) { complete(sends_, rank); }
  void onRecv(int rank) { complete(recvs_, rank); }
  void fail(const std::string& why) {
    std::lock_guard<std::mutex> g(mu_);
    error_ = why;
    cv_.notify_all();
  }
  uint32_t lkey() const { return mr_->mr->lkey; }
  uint32_t rkey() const { return mr_->mr->rkey; }
  bool hostMemory() const { return host_; }
  size_t span(size_t offset, size_t nbytes) const {
    if (nbytes == kUnspecifiedByteCount) {
      GLB_ENFORCE_LE(offset, size);
      nbytes = size - offset;
    }
    GLB_ENFORCE(offset <= size && nbytes <= size - offset, "range exceeds the buffer");
    return nbytes;
  }

 private:
  void complete(std::deque<int>& q, int rank) {
    std::lock_guard<std::mutex> g(mu_);
    q.push_back(rank);
    cv_.notify_all();
  }
  bool wait(std::deque<int>& q, bool& aborted, int* rank, std::chrono::milliseconds timeout, const char* what);
  Ctx* ctx_;
  std::shared_ptr<Mr> mr_;
  bool host_ = true;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<int> sends_, recvs_;
  bool abortSend_ = false, abortRecv_ = false;
  std::string error_;
};

// ---- context: pairs + receiver-local matching of unbound messages ---------------------------------

struct Arrival {
  int src = -1;
  uint64_t slot = 0;
  bool eager = false;
  std::vector<char> data;  // eager payload
  uint64_t raddr = 0;      // rendezvous
  uint32_t rkey = 0;
  uint32_t id = 0;
  size_t nbytes = 0;
};

struct Posted {
  UBuf* buf;
  std::set<int> srcs;
  uint64_t slot;
  size_t offset, nbytes;
};

class Ctx : public Context {
 public:
  Ctx(std::shared_ptr<Dev> dev, int rank, int size) : Context(rank, size), dev_(std::move(dev)) {}
  ~Ctx() override {
    // pairs reference the context: destroy them first
    for (auto& p : pairs_) p.reset();
  }
  std::unique_ptr<Pair>& createPair(int r) override;
  std::unique_ptr<UnboundBuffer> createUnboundBuffer(void* ptr, size_t size) override {
    return std::make_unique<UBuf>(this, ptr, size);
  }
  std::unique_ptr<RemoteKey> deserializeRemoteKey(const std::string& s) override {
    int rk = 0;
    unsigned long long addr = 0, size = 0;
    unsigned key = 0;
    GLB_ENFORCE_EQ(std::sscanf(s.c_str(), "ibv:%d:%llu:%u:%llu", &rk, &addr, &key, &size), 4, "malformed ibverbs remote key");
    return std::make_unique<Key>(rk, addr, key, size);
  }
  Dev& dev() { return *dev_; }
  PairImpl* pairImpl(int r);

  void postRecv(Posted p);
  void onArrival(Arrival a);
  void forget(UBuf* b) {
    std::lock_guard<std::mutex> g(matchMu_);
    posted_.remove_if([b](const Posted& p) { return p.buf == b; });
  }

 private:
  void deliver(const Posted& p, Arrival& a);  // requires no lock
  std::shared_ptr<Dev> dev_;
  std::mutex matchMu_;
  std::list<Posted> posted_;
  std::list<Arrival> unexpected_;
};

// ---- pair: one RC queue pair ---------------------------------------------------------------------

class PairImpl : public Pair {
 public:
  PairImpl(Ctx* ctx, Dev& dev, int peer) : ctx_(ctx), dev_(dev), api_(dev.api()), peer_(peer) {
    ibv_qp_init_attr ia;
    std::memset(&ia, 0, sizeof(ia));
    ia.send_cq = dev_.cq();
    ia.recv_cq = dev_.cq();
    ia.cap.max_send_wr = kSendDepth;
    ia.cap.max_recv_wr = kRecvSlots + 8;
    ia.cap.max_send_sge = 2;
    ia.cap.max_recv_sge = 1;
    ia.cap.max_inline_data = kInline;
    ia.qp_type = IBV_QPT_RC;
    qp_ = api_->create_qp(dev_.pd(), &ia);
    GLB_ENFORCE(qp_ != nullptr, "ibv_create_qp failed: ", std::strerror(errno));
    ibv_qp_attr a;
    std::memset(&a, 0, sizeof(a));
    a.qp_state = IBV_QPS_INIT;
    a.pkey_index = 0;
    a.port_num = static_cast<uint8_t>(dev_.port());
    a.qp_access_flags = IBV_ACCESS_LOCAL_WRITE | IBV_ACCESS_REMOTE_READ | IBV_ACCESS_REMOTE_WRITE;
    GLB_ENFORCE_EQ(api_->modify_qp(qp_, &a, IBV_QP_STATE | IBV_QP_PKEY_INDEX | IBV_QP_PORT | IBV_QP_ACCESS_FLAGS), 0,
                   "QP -> INIT failed");
    // receive ring + eager bounce slots, one registration
    slab_.resize((kRecvSlots + kSendSlots) * (sizeof(Header) + kEagerMax));
    slabMr_ = dev_.reg(slab_.data(), slab_.size(), IBV_ACCESS_LOCAL_WRITE);
    for (int i = 0; i < kRecvSlots; i++) postRecvSlot(i);
    for (int i = 0; i < kSendSlots; i++) freeSendSlots_.push_back(i);
    addr_.b.lid = dev_.lid();
    addr_.b.qpn = qp_->qp_num;
    addr_.b.psn = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(this) >> 4) & 0xffffff;
    addr_.b.mtu = static_cast<uint32_t>(dev_.mtu());
    std::memcpy(addr_.b.gid, dev_.gid().raw, 16);
    dev_.addPair(qp_->qp_num, this);
  }
  ~PairImpl() override { close(); }

  const Address& address() const override { return addr_; }

  void connect(const std::vector<char>& bytes) override {
    Addr::Blob peer;
    GLB_ENFORCE_EQ(bytes.size(), sizeof(peer), "malformed ibverbs address");
    std::memcpy(&peer, bytes.data(), sizeof(peer));
    ibv_qp_attr a;
    std::memset(&a, 0, sizeof(a));
    a.qp_state = IBV_QPS_RTR;
    a.path_mtu = static_cast<ibv_mtu>(std::min<uint32_t>(peer.mtu, static_cast<uint32_t>(dev_.mtu())));
    a.dest_qp_num = peer.qpn;
    a.rq_psn = peer.psn;
    a.max_dest_rd_atomic = 16;
    a.min_rnr_timer = 12;
    a.ah_attr.dlid = peer.lid;
    a.ah_attr.port_num = static_cast<uint8_t>(dev_.port());
    a.ah_attr.is_global = 1;  // works for RoCE and for IB (GRH)
    std::memcpy(a.ah_attr.grh.dgid.raw, peer.gid, 16);
    a.ah_attr.grh.sgid_index = static_cast<uint8_t>(dev_.gidIndex());
    a.ah_attr.grh.hop_limit = 1;
    GLB_ENFORCE_EQ(api_->modify_qp(qp_, &a,
                                   IBV_QP_STATE | IBV_QP_AV | IBV_QP_PATH_MTU | IBV_QP_DEST_QPN | IBV_QP_RQ_PSN |
                                       IBV_QP_MAX_DEST_RD_ATOMIC | IBV_QP_MIN_RNR_TIMER),
                   0, "QP -> RTR failed (rank ", peer_, ")");
    std::memset(&a, 0, sizeof(a));
    a.qp_state = IBV_QPS_RTS;
    a.sq_psn = addr_.b.psn;
    a.timeout = 14;
    a.retry_cnt = 7;
    a.rnr_retry = 7;  // infinite: the receive ring is replenished by the completion thread
    a.max_rd_atomic = 16;
    GLB_ENFORCE_EQ(api_->modify_qp(qp_, &a,
                                   IBV_QP_STATE | IBV_QP_TIMEOUT | IBV_QP_RETRY_CNT | IBV_QP_RNR_RETRY | IBV_QP_SQ_PSN |
                                       IBV_QP_MAX_QP_RD_ATOMIC),
                   0, "QP -> RTS failed (rank ", peer_, ")");
    std::lock_guard<std::mutex> g(mu_);
    connected_ = true;
    cv_.notify_all();
  }

  void close() override {
    ibv_qp* qp = nullptr;
    {
      std::lock_guard<std::mutex> g(mu_);
      qp = qp_;
      qp_ = nullptr;
      connected_ = false;
    }
    if (qp != nullptr) {
      dev_.removePair(qp->qp_num);
      api_->destroy_qp(qp);
    }
  }
  bool isConnected() override {
    std::lock_guard<std::mutex> g(mu_);
    return connected_;
  }
  void setSync(bool, bool) override {}  // completions always come from the device thread

  std::unique_ptr<Buffer> createSendBuffer(int slot, void* ptr, size_t size) override {
    return std::make_unique<BoundBuf>(this, slot, ptr, size, false, dev_.reg(ptr, size, IBV_ACCESS_LOCAL_WRITE));
  }
  std::unique_ptr<Buffer> createRecvBuffer(int slot, void* ptr, size_t size) override {
    auto b = std::make_unique<BoundBuf>(this, slot, ptr, size, true,
                                        dev_.reg(ptr, size, IBV_ACCESS_LOCAL_WRITE | IBV_ACCESS_REMOTE_WRITE));
    {
      std::lock_guard<std::mutex> g(mu_);
      GLB_ENFORCE(recvBufs_.count(slot) == 0, "duplicate recv buffer on slot ", slot);
      recvBufs_[slot] = b.get();
    }
    // tell the peer where writes to this slot land
    Header h{};
    h.type = MSG_MR;
    h.slot = static_cast<uint64_t>(slot);
    h.raddr = reinterpret_cast<uint64_t>(ptr);
    h.rkey = b->rkey();
    h.nbytes = size;
    sendControl(h);
    return b;
  }
  void forgetRecvBuffer(int slot) {
    std::lock_guard<std::mutex> g(mu_);
    recvBufs_.erase(slot);
  }

  void send(UnboundBuffer* buf, uint64_t tag, size_t offset, size_t nbytes) override {
    sendUnbound(static_cast<UBuf*>(buf), tag, offset, nbytes);
  }
  void recv(UnboundBuffer* buf, uint64_t tag, size_t offset, size_t nbytes) override {
    static_cast<UBuf*>(buf)->recv(peer_, tag, offset, nbytes);
  }

  // ---- data path, called by buffers -------------------------------------------------------
  void sendUnbound(UBuf* b, uint64_t slot, size_t offset, size_t nbytes) {
    if (b->hostMemory() && nbytes <= kEagerMax) {
      const int s = takeSendSlot();
      char* bounce = sendSlotPtr(s);
      Header h{};
      h.type = MSG_EAGER;
      h.slot = slot;
      h.nbytes = nbytes;
      std::memcpy(bounce, &h, sizeof(h));
      if (nbytes > 0) std::memcpy(bounce + sizeof(h), static_cast<const char*>(b->ptr) + offset, nbytes);
      auto* op = new Op();
      op->kind = Op::EAGER_SLOT;
      op->sendSlot = s;
      ibv_sge sge{reinterpret_cast<uint64_t>(bounce), static_cast<uint32_t>(sizeof(h) + nbytes), slabMr_->mr->lkey};
      ibv_send_wr wr;
      std::memset(&wr, 0, sizeof(wr));
      wr.wr_id = reinterpret_cast<uint64_t>(op);
      wr.sg_list = &sge;
      wr.num_sge = 1;
      wr.opcode = IBV_WR_SEND;
      wr.send_flags = IBV_SEND_SIGNALED;
      post(wr);
      b->onSend(peer_);  // the payload has been copied: the user buffer is free again
      return;
    }
    uint32_t id;
    {
      std::lock_guard<std::mutex> g(mu_);
      id = nextId_++;
      pendingSends_[id] = b;
    }
    Header h{};
    h.type = MSG_RTS;
    h.id = id;
    h.slot = slot;
    h.nbytes = nbytes;
    h.raddr = reinterpret_cast<uint64_t>(static_cast<char*>(b->ptr) + offset);
    h.rkey = b->rkey();
    sendControl(h);
  }

  // Receiver side of a rendezvous: pull the payload, then FIN.
  void readRemote(UBuf* dst, size_t offset, size_t nbytes, uint64_t raddr, uint32_t rkey, uint32_t id) {
    auto* op = new Op();
    op->kind = Op::READ_FOR_RECV;
    op->ubuf = dst;
    op->id = id;
    postRdma(IBV_WR_RDMA_READ, op, static_cast<char*>(dst->ptr) + offset, nbytes, dst->lkey(), raddr, rkey, 0, false);
  }
  void putRemote(UBuf* src, size_t offset, size_t nbytes, uint64_t raddr, uint32_t rkey) {
    auto* op = new Op();
    op->kind = Op::PUT;
    op->ubuf = src;
    postRdma(IBV_WR_RDMA_WRITE, op, static_cast<char*>(src->ptr) + offset, nbytes, src->lkey(), raddr, rkey, 0, false);
  }
  void getRemote(UBuf* dst, size_t offset, size_t nbytes, uint64_t raddr, uint32_t rkey) {
    auto* op = new Op();
    op->kind = Op::GET;
    op->ubuf = dst;
    postRdma(IBV_WR_RDMA_READ, op, static_cast<char*>(dst->ptr) + offset, nbytes, dst->lkey(), raddr, rkey, 0, false);
  }
  void writeBound(BoundBuf* src, size_t offset, size_t length, size_t roffset) {
    // where does the peer want writes to this slot?
    PeerRegion region;
    {
      std::unique_lock<std::mutex> g(mu_);
      const auto deadline = std::chrono::steady_clock::now() + effectiveTimeout();
      while (peerRegions_.count(src->slot()) == 0) {
        throwIfFailed();
        if (cv_.wait_until(g, deadline) == std::cv_status::timeout) {
          GLB_THROW_TIMEOUT("timed out waiting for rank ", peer_, " to create its receive buffer on slot ", src->slot());
        }
      }
      region = peerRegions_[src->slot()];
    }
    GLB_ENFORCE(roffset <= region.size && length <= region.size - roffset, "bound write exceeds the peer's buffer");
    auto* op = new Op();
    op->kind = Op::BOUND_WRITE;
    op->bound = src;
    postRdma(IBV_WR_RDMA_WRITE_WITH_IMM, op, static_cast<char*>(src->ptr()) + offset, length, src->lkey(),
             region.addr + roffset, region.rkey, static_cast<uint32_t>(src->slot()), true);
  }

  // ---- completion thread --------------------------------------------------------------------
  void handleCompletion(const ibv_wc& wc) {
    if (wc.status != IBV_WC_SUCCESS) {
      failAll(strcat_all("work completion error ", static_cast<int>(wc.status), " (vendor ", wc.vendor_err, ") on the pair to rank ", peer_));
      // only wr_id / status / qp_num are valid in a failed completion; receive work requests
      // carry their ring index, send-side ones a pointer
      if (wc.wr_id > 0xffff) delete reinterpret_cast<Op*>(wc.wr_id);
      return;
    }
    if (wc.opcode & IBV_WC_RECV) {
      const int s = static_cast<int>(wc.wr_id);
      if (wc.opcode == IBV_WC_RECV_RDMA_WITH_IMM) {
        BoundBuf* b = nullptr;
        {
          std::lock_guard<std::mutex> g(mu_);
          auto it = recvBufs_.find(static_cast<int>(ntohl(wc.imm_data)));
          if (it != recvBufs_.end()) b = it->second;
        }
        if (b != nullptr) b->onRecv();
      } else {
        handleMessage(recvSlotPtr(s), wc.byte_len);
      }
      postRecvSlot(s);
      return;
    }
    std::unique_ptr<Op> op(reinterpret_cast<Op*>(wc.wr_id));
    {
      std::lock_guard<std::mutex> g(mu_);
      outstanding_--;
      cv_.notify_all();
    }
    switch (op->kind) {
      case Op::CONTROL: break;
      case Op::EAGER_SLOT: releaseSendSlot(op->sendSlot); break;
      case Op::BOUND_WRITE: op->bound->onSend(); break;
      case Op::PUT: op->ubuf->onSend(peer_); break;
      case Op::GET: op->ubuf->onRecv(peer_); break;
      case Op::READ_FOR_RECV: {
        op->ubuf->onRecv(peer_);
        Header h{};
        h.type = MSG_FIN;
        h.id = op->id;
        sendControl(h, /*fromCompletionThread=*/true);
        break;
      }
    }
    flushDeferred();
  }

  void failAll(const std::string& why) {
    std::map<uint32_t, UBuf*> sends;
    std::map<int, BoundBuf*> recvs;
    {
      std::lock_guard<std::mutex> g(mu_);
      if (!error_.empty()) return;
      error_ = why;
      sends.swap(pendingSends_);
      recvs = recvBufs_;
      cv_.notify_all();
    }
    GLB_ERROR("ibverbs: ", why);
    for (auto& kv : sends) kv.second->fail(why);
    for (auto& kv : recvs) kv.second->fail(why);
  }

  int peer() const { return peer_; }
  std::chrono::milliseconds effectiveTimeout() const;

 private:
  struct PeerRegion {
    uint64_t addr = 0;
    uint32_t rkey = 0;
    size_t size = 0;
  };

  char* recvSlotPtr(int s) { return slab_.data() + static_cast<size_t>(s) * (sizeof(Header) + kEagerMax); }
  char* sendSlotPtr(int s) { return slab_.data() + static_cast<size_t>(kRecvSlots + s) * (sizeof(Header) + kEagerMax); }

  void postRecvSlot(int s) {
    ibv_sge sge{reinterpret_cast<uint64_t>(recvSlotPtr(s)), static_cast<uint32_t>(sizeof(Header) + kEagerMax), slabMr_->mr->lkey};
    ibv_recv_wr wr;
    std::memset(&wr, 0, sizeof(wr));
    wr.wr_id = static_cast<uint64_t>(s);
    wr.sg_list = &sge;
    wr.num_sge = 1;
    ibv_recv_wr* bad = nullptr;
    ibv_qp* qp = qp_;
    if (qp != nullptr && dev_.ctx()->ops.post_recv(qp, &wr, &bad) != 0) failAll("ibv_post_recv failed");
  }

  int takeSendSlot() {
    std::unique_lock<std::mutex> g(mu_);
    const auto deadline = std::chrono::steady_clock::now() + effectiveTimeout();
    while (freeSendSlots_.empty()) {
      throwIfFailed();
      if (cv_.wait_until(g, deadline) == std::cv_status::timeout) GLB_THROW_TIMEOUT("no eager slot became free (rank ", peer_, ")");
    }
    const int s = freeSendSlots_.back();
    freeSendSlots_.pop_back();
    return s;
  }
  void releaseSendSlot(int s) {
    std::lock_guard<std::mutex> g(mu_);
    freeSendSlots_.push_back(s);
    cv_.notify_all();
  }

  void throwIfFailed() {  // requires mu_
    if (!error_.empty()) GLB_THROW_IO_EXCEPTION(error_);
  }

  // Header-only message, sent inline. From the completion thread it must never block: when
  // the send queue is full the message is parked and flushed by later completions.
  void sendControl(const Header& h, bool fromCompletionThread = false) {
    {
      std::unique_lock<std::mutex> g(mu_);
      if (fromCompletionThread) {
        if (outstanding_ >= kSendDepth - 4) {
          deferred_.push_back(h);
          return;
        }
      } else {
        const auto deadline = std::chrono::steady_clock::now() + effectiveTimeout();
        while (!connected_ || outstanding_ >= kSendDepth - 16) {
          throwIfFailed();
          if (cv_.wait_until(g, deadline) == std::cv_status::timeout) GLB_THROW_TIMEOUT("send queue to rank ", peer_, " stayed full");
        }
      }
    }
    Header copy = h;
    auto* op = new Op();
    op->kind = Op::CONTROL;
    ibv_sge sge{reinterpret_cast<uint64_t>(&copy), static_cast<uint32_t>(sizeof(copy)), 0};
    ibv_send_wr wr;
    std::memset(&wr, 0, sizeof(wr));
    wr.wr_id = reinterpret_cast<uint64_t>(op);
    wr.sg_list = &sge;
    wr.num_sge = 1;
    wr.opcode = IBV_WR_SEND;
    wr.send_flags = IBV_SEND_SIGNALED | IBV_SEND_INLINE;
    post(wr);
  }
  void flushDeferred() {
    while (true) {
      Header h;
      {
        std::lock_guard<std::mutex> g(mu_);
        if (deferred_.empty() || outstanding_ >= kSendDepth - 4) return;
        h = deferred_.front();
        deferred_.pop_front();
      }
      sendControl(h, true);
    }
  }

  void postRdma(ibv_wr_opcode opcode, Op* op, void* local, size_t nbytes, uint32_t lkey, uint64_t raddr, uint32_t rkey,
                uint32_t imm, bool withImm) {
    {
      std::unique_lock<std::mutex> g(mu_);
      const auto deadline = std::chrono::steady_clock::now() + effectiveTimeout();
      while (!connected_ || outstanding_ >= kSendDepth - 16) {
        if (!error_.empty()) {
          delete op;
          throwIfFailed();
        }
        if (cv_.wait_until(g, deadline) == std::cv_status::timeout) {
          delete op;
          GLB_THROW_TIMEOUT("send queue to rank ", peer_, " stayed full");
        }
      }
    }
    // One work request moves at most 2 GiB - 1: larger transfers are cut; only the last piece is signaled.
    constexpr size_t kMaxWr = 1u << 30;
    size_t done = 0;
    do {
      const size_t n = std::min(kMaxWr, nbytes - done);
      const bool last = done + n >= nbytes;
      ibv_sge sge{reinterpret_cast<uint64_t>(static_cast<char*>(local) + done), static_cast<uint32_t>(n), lkey};
      ibv_send_wr wr;
      std::memset(&wr, 0, sizeof(wr));
      wr.sg_list = &sge;
      wr.num_sge = n > 0 ? 1 : 0;
      wr.wr.rdma.remote_addr = raddr + done;
      wr.wr.rdma.rkey = rkey;
      if (last) {
        wr.wr_id = reinterpret_cast<uint64_t>(op);
        wr.opcode = opcode;
        wr.send_flags = IBV_SEND_SIGNALED;
        if (withImm) wr.imm_data = htonl(imm);
        post(wr);
      } else {
        auto* filler = new Op();
        filler->kind = Op::CONTROL;
        wr.wr_id = reinterpret_cast<uint64_t>(filler);
        wr.opcode = opcode == IBV_WR_RDMA_WRITE_WITH_IMM ? IBV_WR_RDMA_WRITE : opcode;
        wr.send_flags = IBV_SEND_SIGNALED;
        post(wr);
      }
      done += n;
    } while (done < nbytes);
  }

  void post(ibv_send_wr& wr) {
    ibv_send_wr* bad = nullptr;
    ibv_qp* qp;
    {
      std::lock_guard<std::mutex> g(mu_);
      qp = qp_;
      outstanding_++;
    }
    if (qp == nullptr || dev_.ctx()->ops.post_send(qp, &wr, &bad) != 0) {
      {
        std::lock_guard<std::mutex> g(mu_);
        outstanding_--;
      }
      delete reinterpret_cast<Op*>(wr.wr_id);
      failAll(strcat_all("ibv_post_send to rank ", peer_, " failed"));
      GLB_THROW_IO_EXCEPTION("ibv_post_send to rank ", peer_, " failed");
    }
  }

  void handleMessage(const char* slotData, uint32_t len) {
    if (len < sizeof(Header)) return failAll("short message");
    Header h;
    std::memcpy(&h, slotData, sizeof(h));
    switch (h.type) {
      case MSG_EAGER: {
        if (h.nbytes > kEagerMax || sizeof(Header) + h.nbytes > len) return failAll("malformed eager message");
        Arrival a;
        a.src = peer_;
        a.slot = h.slot;
        a.eager = true;
        a.nbytes = h.nbytes;
        a.data.assign(slotData + sizeof(Header), slotData + sizeof(Header) + h.nbytes);
        ctx_->onArrival(std::move(a));
        break;
      }
      case MSG_RTS: {
        Arrival a;
        a.src = peer_;
        a.slot = h.slot;
        a.nbytes = h.nbytes;
        a.raddr = h.raddr;
        a.rkey = h.rkey;
        a.id = h.id;
        ctx_->onArrival(std::move(a));
        break;
      }
      case MSG_FIN: {
        UBuf* b = nullptr;
        {
          std::lock_guard<std::mutex> g(mu_);
          auto it = pendingSends_.find(h.id);
          if (it != pendingSends_.end()) {
            b = it->second;
            pendingSends_.erase(it);
          }
        }
        if (b != nullptr) b->onSend(peer_);
        break;
      }
      case MSG_MR: {
        std::lock_guard<std::mutex> g(mu_);
        peerRegions_[static_cast<int>(h.slot)] = PeerRegion{h.raddr, h.rkey, static_cast<size_t>(h.nbytes)};
        cv_.notify_all();
        break;
      }
      default: failAll("unknown message type");
    }
  }

 public:
  void forgetSend(UBuf* b) {
    std::lock_guard<std::mutex> g(mu_);
    for (auto it = pendingSends_.begin(); it != pendingSends_.end();) it = it->second == b ? pendingSends_.erase(it) : std::next(it);
  }

 private:
  Ctx* ctx_;
  Dev& dev_;
  const VerbsApi* api_;
  int peer_;
  ibv_qp* qp_ = nullptr;
  Addr addr_;
  std::vector<char> slab_;
  std::shared_ptr<Mr> slabMr_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool connected_ = false;
  int outstanding_ = 0;
  std::vector<int> freeSendSlots_;
  std::deque<Header> deferred_;
  std::map<int, BoundBuf*> recvBufs_;
  std::map<int, PeerRegion> peerRegions_;
  std::map<uint32_t, UBuf*> pendingSends_;
  uint32_t nextId_ = 1;
  std::string error_;
};

std::chrono::milliseconds PairImpl::effectiveTimeout() const {
  auto t = ctx_->getTimeout();
  return t.count() > 0 ? t : std::chrono::milliseconds(std::chrono::hours(24 * 365));
}

// ---- Dev: completion thread ----------------------------------------------------------------------

void Dev::drain() {
  ibv_wc wc[32];
  while (true) {
    const int n = ctx_->ops.poll_cq(cq_, 32, wc);
    if (n <= 0) return;
    for (int i = 0; i < n; i++) {
      PairImpl* p = nullptr;
      {
        std::lock_guard<std::mutex> g(mu_);
        auto it = pairs_.find(wc[i].qp_num);
        if (it != pairs_.end()) {
          p = it->second;
          busyQpn_ = wc[i].qp_num;
        }
      }
      if (p != nullptr) {
        try {
          p->handleCompletion(wc[i]);
        } catch (const std::exception& e) {
          p->failAll(e.what());
        }
        std::lock_guard<std::mutex> g(mu_);
        busyQpn_ = 0;
      } else if (wc[i].wr_id > 0xffff) {
        delete reinterpret_cast<Op*>(wc[i].wr_id);  // completion of a pair that is already gone
      }
    }
  }
}

void Dev::loop() {
  while (!done_) {
    struct pollfd pfd = {channel_->fd, POLLIN, 0};
    const int rv = ::poll(&pfd, 1, 50);
    if (rv > 0) {
      ibv_cq* cq = nullptr;
      void* cqctx = nullptr;
      if (api_->get_cq_event(channel_, &cq, &cqctx) == 0) {
        api_->ack_cq_events(cq, 1);
        ctx_->ops.req_notify_cq(cq_, 0);
      }
    }
    drain();
  }
}

std::shared_ptr<Context> Dev::createContext(int rank, int size) { return std::make_shared<Ctx>(shared_from_this(), rank, size); }

// ---- Ctx ---------------------------------------------------------------------------------------

std::unique_ptr<Pair>& Ctx::createPair(int r) {
  pairs_[r] = std::make_unique<PairImpl>(this, *dev_, r);
  return pairs_[r];
}

PairImpl* Ctx::pairImpl(int r) {
  GLB_ENFORCE(r >= 0 && r < size && r != rank, "invalid peer rank ", r);
  auto* p = static_cast<PairImpl*>(pairs_[r].get());
  GLB_ENFORCE(p != nullptr, "no pair to rank ", r);
  return p;
}

void Ctx::deliver(const Posted& p, Arrival& a) {
  if (a.nbytes > p.nbytes) {
    p.buf->fail(strcat_all("message of ", a.nbytes, " bytes from rank ", a.src, " does not fit the posted receive of ", p.nbytes));
    return;
  }
  if (a.eager) {
    if (!p.buf->hostMemory()) {
      p.buf->fail("eager message for a device buffer");  // senders use rendezvous for device memory on both ends
      return;
    }
    if (a.nbytes > 0) std::memcpy(static_cast<char*>(p.buf->ptr) + p.offset, a.data.data(), a.nbytes);
    p.buf->onRecv(a.src);
  } else {
    pairImpl(a.src)->readRemote(p.buf, p.offset, a.nbytes, a.raddr, a.rkey, a.id);
  }
}

void Ctx::postRecv(Posted p) {
  Arrival hit;
  bool found = false;
  {
    std::lock_guard<std::mutex> g(matchMu_);
    for (auto it = unexpected_.begin(); it != unexpected_.end(); ++it) {
      if (it->slot == p.slot && p.srcs.count(it->src)) {
        hit = std::move(*it);
        unexpected_.erase(it);
        found = true;
        break;
      }
    }
    if (!found) {
      posted_.push_back(std::move(p));
      return;
    }
  }
  deliver(p, hit);
}

void Ctx::onArrival(Arrival a) {
  Posted hit{};
  bool found = false;
  {
    std::lock_guard<std::mutex> g(matchMu_);
    for (auto it = posted_.begin(); it != posted_.end(); ++it) {
      if (it->slot == a.slot && it->srcs.count(a.src)) {
        hit = *it;
        posted_.erase(it);
        found = true;
        break;
      }
    }
    if (!found) {
      unexpected_.push_back(std::move(a));
      return;
    }
  }
  deliver(hit, a);
}

// ---- BoundBuf / UBuf ---------------------------------------------------------------------------

BoundBuf::~BoundBuf() {
  if (isRecv_) pair_->forgetRecvBuffer(slot_);
}

void BoundBuf::send(size_t offset, size_t length, size_t roffset) {
  GLB_ENFORCE(!isRecv_, "send() on a receive buffer");
  GLB_ENFORCE(offset <= size_ && length <= size_ - offset, "bound send exceeds the buffer");
  pair_->writeBound(this, offset, length, roffset);
}

void BoundBuf::wait(int& counter, const char* what) {
  std::unique_lock<std::mutex> g(mu_);
  const auto deadline = std::chrono::steady_clock::now() + pair_->effectiveTimeout();
  while (counter == 0) {
    if (!error_.empty()) GLB_THROW_IO_EXCEPTION(error_);
    if (cv_.wait_until(g, deadline) == std::cv_status::timeout) {
      GLB_THROW_TIMEOUT("timed out waiting for a bound ", what, " on slot ", slot_, " (rank ", pair_->peer(), ")");
    }
  }
  counter--;
}

// Device memory cannot be copied through a host bounce slot: such buffers always use the
// rendezvous (RDMA READ) path.
bool isHostPointer(const void* p) { return cuda::deviceCount() <= 0 || cuda::deviceForPointer(p) < 0; }

UBuf::UBuf(Ctx* ctx, void* ptr, size_t size) : UnboundBuffer(ptr, size), ctx_(ctx) {
  host_ = isHostPointer(ptr);
  mr_ = ctx_->dev().reg(ptr, size, IBV_ACCESS_LOCAL_WRITE | IBV_ACCESS_REMOTE_READ | IBV_ACCESS_REMOTE_WRITE);
}

UBuf::~UBuf() {
  ctx_->forget(this);
  for (int r = 0; r < ctx_->size; r++) {
    if (r == ctx_->rank) continue;
    if (auto* p = static_cast<PairImpl*>(ctx_->peekPair(r))) p->forgetSend(this);
  }
}

bool UBuf::wait(std::deque<int>& q, bool& aborted, int* rank, std::chrono::milliseconds timeout, const char* what) {
  if (timeout == kUnsetTimeout) timeout = ctx_->getTimeout();
  std::unique_lock<std::mutex> g(mu_);
  const bool bounded = timeout != kNoTimeout && timeout.count() > 0;
  const auto deadline = std::chrono::steady_clock::now() + (bounded ? timeout : std::chrono::milliseconds(0));
  while (q.empty()) {
    if (!error_.empty()) GLB_THROW_IO_EXCEPTION(error_);
    if (aborted) {
      aborted = false;
      return false;
    }
    if (bounded) {
      if (cv_.wait_until(g, deadline) == std::cv_status::timeout && q.empty()) {
        GLB_THROW_TIMEOUT("timed out waiting for an unbound ", what, " after ", timeout.count(), " ms");
      }
    } else {
      cv_.wait(g);
    }
  }
  if (rank != nullptr) *rank = q.front();
  q.pop_front();
  return true;
}

void UBuf::send(int dstRank, uint64_t slot, size_t offset, size_t nbytes) {
  nbytes = span(offset, nbytes);
  ctx_->pairImpl(dstRank)->sendUnbound(this, slot, offset, nbytes);
}

void UBuf::recv(std::vector<int> srcRanks, uint64_t slot, size_t offset, size_t nbytes) {
  nbytes = span(offset, nbytes);
  Posted p{this, std::set<int>(srcRanks.begin(), srcRanks.end()), slot, offset, nbytes};
  ctx_->postRecv(std::move(p));
}

std::unique_ptr<RemoteKey> UBuf::getRemoteKey() const {
  return std::make_unique<Key>(ctx_->rank, reinterpret_cast<uint64_t>(ptr), rkey(), size);
}

void UBuf::put(const RemoteKey& key, uint64_t /*slot*/, size_t offset, size_t roffset, size_t nbytes) {
  const auto* k = dynamic_cast<const Key*>(&key);
  GLB_ENFORCE(k != nullptr, "put() needs an ibverbs remote key");
  GLB_ENFORCE(offset <= size && nbytes <= size - offset, "put: local range exceeds the buffer");
  GLB_ENFORCE(roffset <= k->size && nbytes <= k->size - roffset, "put: remote range exceeds the region of the key");
  ctx_->pairImpl(k->rank)->putRemote(this, offset, nbytes, k->addr + roffset, k->rkey);
}

void UBuf::get(const RemoteKey& key, uint64_t /*slot*/, size_t offset, size_t roffset, size_t nbytes) {
  const auto* k = dynamic_cast<const Key*>(&key);
  GLB_ENFORCE(k != nullptr, "get() needs an ibverbs remote key");
  GLB_ENFORCE(offset <= size && nbytes <= size - offset, "get: local range exceeds the buffer");
  GLB_ENFORCE(roffset <= k->size && nbytes <= k->size - roffset, "get: remote range exceeds the region of the key");
  ctx_->pairImpl(k->rank)->getRemote(this, offset, nbytes, k->addr + roffset, k->rkey);
}

}  // namespace

// ---- public entry points ------------------------------------------------------------------------

Probe probe() {
  Probe p;
  const auto& mods = kernelModules();
  p.peerMemoryModule = mods.count("nv_peer_mem") > 0 || mods.count("nvidia_peermem") > 0;
  std::string why;
  const VerbsApi* api = verbs(&why);
  if (api == nullptr) {
    p.detail = why;
    return p;
  }
  p.libraryLoaded = true;
  int n = 0;
  ibv_device** list = api->get_device_list(&n);
  if (list != nullptr) {
    for (int i = 0; i < n; i++) {
      const char* name = api->get_device_name(list[i]);
      if (name != nullptr) p.devices.emplace_back(name);
    }
    api->free_device_list(list);
  }
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
                                          "CUDA buffers, transport::nvl / cuda::PeerContext over NVLink.");
  }
  return std::make_shared<Dev>(a, verbs());
}

}  // namespace ibverbs
}  // namespace transport
}  // namespace glb
