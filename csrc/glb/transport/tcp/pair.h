// tcp::Pair — one TCP connection to one peer, multiplexing every slot.
//
// Wire protocol (little-endian, 48-byte header followed by `nbytes` of payload):
//
//   SEND_UNBOUND  eager message for an UnboundBuffer recv posted on `slot`. If no
//                 matching recv is posted yet the payload is parked in the
//                 context's unexpected queue and copied when the recv arrives —
//                 so there is no sender/receiver handshake on the critical path
//                 (the reference needs NOTIFY_SEND_READY / NOTIFY_RECV_READY round
//                 trips, pair.cc:913-975), and recv-from-any is a purely local
//                 match on the receiver.
//   SEND_BOUND    one-sided write into the peer's bound recv Buffer registered
//                 under `slot`, at `roffset`. Arrivals that precede registration
//                 are parked and applied by createRecvBuffer().
//   PUT / GET_REQ / GET_RESP
//                 software one-sided access to a region exported through
//                 UnboundBuffer::getRemoteKey() — served entirely by the remote
//                 I/O thread, the remote user thread is not involved.
//   CAPS / CAPS_OK / FIN
//                 same-host single-copy path. Each side announces (pid, address of a
//                 per-process random probe word). A receiver that can read that word
//                 with process_vm_readv answers CAPS_OK; from then on the peer sends
//                 payloads of >= GLB_TCP_CMA_MIN bytes as a header only (flag F_CMA,
//                 `length` = source address) and the receiver pulls the bytes straight
//                 from the sender's address space into the destination — one copy, no
//                 socket buffers — then answers FIN, which completes the send. Pulls
//                 happen where socket reads would (loop thread or spinning waiter), so
//                 a matched or bound message never depends on the receiving user thread.
//                 An unbound message whose recv is not posted yet is parked as a
//                 descriptor and pulled by the thread that posts the recv (no staging
//                 copy); its sender completes then - the reference's rendezvous
//                 semantics, which its eager small messages do not have. FIN carries the
//                 id the sender put in the header.
//
// Threading: all pair state is guarded by `mu_`. In async mode the device loop
// thread performs reads and flushes queued writes; writes are attempted inline
// from the calling thread first. In sync mode the socket is detached from epoll
// and the waiting user thread drives reads itself (optionally busy-polling).
// Parity: gloo/transport/tcp/pair.{h,cc}.
#pragma once

#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "glb/common/memory.h"
#include "glb/transport/pair.h"
#include "glb/transport/tcp/address.h"
#include "glb/transport/tcp/device.h"
#include "glb/transport/tcp/loop.h"

namespace glb {
namespace transport {
namespace tcp {

class Buffer;
class Context;
class UnboundBuffer;

struct WireHeader {
  static constexpr uint32_t kMagic = 0x4d424c47;  // "GLBM"
  uint32_t magic = kMagic;
  uint16_t opcode = 0;
  uint16_t flags = 0;
  uint64_t slot = 0;     // unbound/bound slot, or request id for GET_*
  uint64_t nbytes = 0;   // payload bytes that follow
  uint64_t roffset = 0;  // destination offset (bound / put) or source offset (get)
  uint64_t aux = 0;      // region id for PUT / GET_REQ
  uint64_t length = 0;   // GET_REQ: bytes requested
};
static_assert(sizeof(WireHeader) == 48, "wire header must be 48 bytes");

enum Opcode : uint16_t {
  OP_SEND_UNBOUND = 1,
  OP_SEND_BOUND = 2,
  OP_PUT = 3,
  OP_GET_REQ = 4,
  OP_GET_RESP = 5,
  OP_CAPS = 6,
  OP_CAPS_OK = 7,
  OP_FIN = 8,
};

enum WireFlags : uint16_t {
  F_ERROR = 0x1,  // GET_RESP: remote lookup failed
  F_CMA = 0x2,    // payload is not inline; pull `nbytes` from `length` in the sender's address space
};

class Pair : public ::glb::transport::Pair, private Handler {
 public:
  enum State { INITIALIZING = 1, CONNECTING = 2, CONNECTED = 3, CLOSED = 4 };

  Pair(Context* context, Device* device, int selfRank, int peerRank, std::chrono::milliseconds timeout,
       bool lazy);
  ~Pair() override;

  const Address& address() const override { return self_; }
  void connect(const std::vector<char>& bytes) override;
  void close() override;
  bool isConnected() override;
  void setSync(bool sync, bool busyPoll) override;

  std::unique_ptr<::glb::transport::Buffer> createSendBuffer(int slot, void* ptr, size_t size) override;
  std::unique_ptr<::glb::transport::Buffer> createRecvBuffer(int slot, void* ptr, size_t size) override;

  void send(::glb::transport::UnboundBuffer* buf, uint64_t tag, size_t offset, size_t nbytes) override;
  void recv(::glb::transport::UnboundBuffer* buf, uint64_t tag, size_t offset, size_t nbytes) override;

  int peerRank() const { return peerRank_; }
  std::chrono::milliseconds timeout() const { return timeout_; }
  bool isSync() const { return sync_; }
  bool isBusyPoll() const { return busyPoll_; }

  // Dial / wait for the inbound connection if that has not happened yet (lazy mode).
  // `setup`: called while the mesh is being built - a peer that connected and has already
  // finished and closed again (it was quick, we were slow) is not a failure of the setup.
  void ensureConnected(bool setup = false);

  // ---- used by Buffer / UnboundBuffer / Context -------------------------------
  void sendBound(Buffer* buf, size_t offset, size_t length, size_t roffset);
  void sendUnbound(UnboundBuffer* buf, uint64_t slot, size_t offset, size_t nbytes);
  void sendPut(UnboundBuffer* buf, uint64_t regionId, size_t offset, size_t roffset, size_t nbytes);
  void sendGetRequest(uint64_t requestId, uint64_t regionId, size_t roffset, size_t nbytes);
  void unregisterBuffer(Buffer* buf);
  // Drop queued sends that reference `buf` (it is being destroyed).
  void forgetUnbound(UnboundBuffer* buf);

  // A recv was posted for a single-copy message that had been parked: fetch `nbytes` from
  // `srcAddr` in the peer's address space into `dst` and release the sender (FIN `id`).
  // Called from the posting thread; throws IoException if the bytes cannot be read.
  void pullDeferred(char* dst, uint64_t srcAddr, size_t nbytes, uint64_t id);

  // Poison this pair: all pending and future operations throw IoException(msg).
  void signalExceptionExternal(const std::string& msg);
  // In sync mode: drive the socket from the calling thread until pred() holds.
  // `lock` must hold mu() on entry and holds it on return.
  void syncWait(std::unique_lock<std::mutex>& lock, const std::function<bool()>& pred,
                std::chrono::milliseconds timeout, const char* what);
  // Async mode, spin-then-block: for a bounded time (GLB_TCP_SPIN_US, default 1000,
  // 0 disables) the waiting user thread pulls bytes off the socket itself instead of
  // sleeping on a condvar until the loop thread has done so. That removes two thread
  // wake-ups (epoll thread, then waiter) from the small-message critical path while
  // the loop thread remains the fallback for everything the waiter does not get to.
  // `lock` must hold mu() on entry and holds it on return.
  void spinWait(std::unique_lock<std::mutex>& lock, const std::function<bool()>& pred);
  // Same, for waiters that do not own mu_: one non-blocking read pass if the pair
  // mutex is free. Never throws, never blocks.
  void tryProgress();
  static int64_t spinBudgetNanos();
  // Parks this pair's descriptor in epoll for the lifetime of the guard (GLB_TCP_PARK_EPOLL=0
  // disables). `locked`: the caller holds mu() when the guard is created and when it dies.
  class PauseGuard {
   public:
    PauseGuard(Pair* p, bool locked);
    ~PauseGuard();
    PauseGuard(const PauseGuard&) = delete;
    PauseGuard& operator=(const PauseGuard&) = delete;

   private:
    Pair* pair_;
    bool locked_;
    bool took_ = false;
  };
  // Process-wide counters of the single-copy path (tests and diagnostics).
  static uint64_t cmaMessages();
  static uint64_t cmaBytes();
  std::mutex& mu() { return mu_; }
  void throwIfException();  // requires mu_

 protected:
  // I/O primitives; the TLS pair overrides these.
  virtual ssize_t ioRecv(void* buf, size_t len);
  virtual ssize_t ioSend(const struct iovec* iov, int iovcnt);
  virtual void ioHandshake(bool isInitiator) {}
  virtual void ioShutdown() {}
  // Bytes already buffered above the socket (TLS): the read loop must not yield to
  // epoll while this is true, the fd would not become readable again.
  virtual bool ioPending() { return false; }
  // Whether payloads may bypass the socket on the same host (the TLS pair says no).
  virtual bool allowCma() const { return true; }

  int fd() const { return fd_; }
  // Detach from the loop thread and wait until it can no longer be inside this pair.
  // Subclasses that own state the I/O hooks use (the TLS session) call this first thing in
  // their destructor - the base destructor would only get to it after they are gone.
  void quiesce();

 private:
  struct TxOp {
    WireHeader hdr;
    const char* data = nullptr;
    size_t nbytes = 0;
    size_t sent = 0;
    bool hasUbuf = false;
    bool notify = true;
    bool bestEffort = false;  // capability chatter: a peer that is already gone is not an error
    bool cma = false;  // header-only on the wire; completion on FIN
    uint64_t cmaId = 0;
    WeakAnchor<UnboundBuffer> ubuf;
    UnboundBuffer* ubufRaw = nullptr;
    Buffer* bbuf = nullptr;
  };

  enum RxKind {
    RX_NONE,
    RX_UNBOUND_DIRECT,
    RX_UNBOUND_UNEXPECTED,
    RX_BOUND_DIRECT,
    RX_BOUND_UNEXPECTED,
    RX_PUT,
    RX_GET_RESP,
    RX_DISCARD
  };
  struct Rx {
    WireHeader hdr;
    size_t hdrRead = 0;
    size_t payloadRead = 0;
    RxKind kind = RX_NONE;
    char* dst = nullptr;
    Lease<UnboundBuffer> ubuf;
    Buffer* bbuf = nullptr;
    bool deferred = false;  // single-copy message parked for its recv: nothing to read, no FIN yet
    std::vector<char> stash;
    void reset() {
      hdrRead = 0;
      payloadRead = 0;
      kind = RX_NONE;
      dst = nullptr;
      ubuf.release();
      bbuf = nullptr;
      deferred = false;
      std::vector<char>().swap(stash);
    }
  };
  struct ParkedBound {
    size_t roffset;
    std::vector<char> data;
  };

  void handleEvents(int events) override;

  void dial();                                     // initiator side
  void attachSocket(Socket sock, bool initiator);  // both sides
  void waitUntilConnected(std::unique_lock<std::mutex>& lock, bool setup);

  void enqueue(TxOp&& op);      // requires mu_
  bool tryWrite(TxOp& op);      // requires mu_; true when fully written
  void completeTx(TxOp& op);    // requires mu_
  void wroteTx(TxOp&& op);      // requires mu_; completes now, or parks until FIN
  void setPayload(TxOp& op, const char* data, size_t nbytes);  // requires mu_
  bool pullPayload();           // requires mu_; false after signalException
  void flushTx();               // requires mu_
  void readLoop(size_t budget); // requires mu_
  void beginMessage();          // requires mu_
  void finishMessage();         // requires mu_
  void signalException(const std::string& msg);  // requires mu_
  void armEvents(bool wantWrite);                // requires mu_
  bool pauseEventsLocked();                      // requires mu_; false if parking is not possible now
  void resumeEventsLocked();                     // requires mu_

  Context* const context_;
  Device* const device_;
  const int selfRank_;
  const int peerRank_;
  const std::chrono::milliseconds timeout_;
  const bool lazy_;
  Loop* loop_;

  std::mutex dialMu_;
  std::mutex mu_;
  std::condition_variable cv_;
  State state_ = INITIALIZING;
  bool sync_ = false;
  bool busyPoll_ = false;
  bool wantWrite_ = false;
  int paused_ = 0;  // waiters currently polling the socket themselves (descriptor parked)
  bool expecting_ = false;
  int fd_ = -1;
  Address self_;
  Address peer_;
  bool havePeer_ = false;
  std::string exMsg_;
  bool failed_ = false;
  bool quiesced_ = false;
  bool everConnected_ = false;  // reached CONNECTED at some point
  bool closedByPeer_ = false;   // the failure is the peer's orderly (or abortive) close

  std::deque<TxOp> tx_;
  std::deque<TxOp> awaitingFin_;  // CMA sends written to the wire, completed by FIN(id)
  uint64_t cmaSeq_ = 0;
  bool copyFrom(char* dst, uint64_t srcAddr, size_t nbytes, std::string* err);  // no lock needed
  bool peerCanPull_ = false;      // peer answered CAPS_OK
  bool canPull_ = false;          // we can read the peer's memory
  int peerPid_ = -1;
  Rx rx_;
  std::unordered_map<int, Buffer*> recvBuffers_;
  std::unordered_map<int, std::deque<ParkedBound>> parkedBound_;
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
