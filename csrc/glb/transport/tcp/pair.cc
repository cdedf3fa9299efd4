#include "glb/transport/tcp/pair.h"

#include <cstdio>
#include <cstdlib>

#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sched.h>
#include <sys/ioctl.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <thread>

#include "glb/common/logging.h"
#include "glb/common/utils.h"
#include "glb/transport/tcp/buffer.h"
#include "glb/transport/tcp/context.h"
#include "glb/transport/tcp/unbound_buffer.h"

namespace glb {
namespace transport {
namespace tcp {

namespace {
// offset + length <= size, without the 64-bit wrap-around a hostile header could provoke.
inline bool rangeFits(uint64_t offset, uint64_t length, uint64_t size) { return offset <= size && length <= size - offset; }

// True when the connected socket's two ends carry the same IP address, or the peer is loopback.
bool peerIsThisHost(int fd) {
  struct sockaddr_storage a, b;
  socklen_t la = sizeof(a), lb = sizeof(b);
  if (fd < 0 || ::getsockname(fd, reinterpret_cast<struct sockaddr*>(&a), &la) != 0 ||
      ::getpeername(fd, reinterpret_cast<struct sockaddr*>(&b), &lb) != 0 || a.ss_family != b.ss_family) {
    return false;
  }
  if (a.ss_family == AF_INET) {
    const auto* x = reinterpret_cast<const struct sockaddr_in*>(&a);
    const auto* y = reinterpret_cast<const struct sockaddr_in*>(&b);
    return x->sin_addr.s_addr == y->sin_addr.s_addr || (ntohl(y->sin_addr.s_addr) >> 24) == 127;
  }
  if (a.ss_family == AF_INET6) {
    const auto* x = reinterpret_cast<const struct sockaddr_in6*>(&a);
    const auto* y = reinterpret_cast<const struct sockaddr_in6*>(&b);
    return std::memcmp(&x->sin6_addr, &y->sin6_addr, sizeof(x->sin6_addr)) == 0 || IN6_IS_ADDR_LOOPBACK(&y->sin6_addr);
  }
  return a.ss_family == AF_UNIX;
}
}  // namespace

namespace {
constexpr size_t kReadBudget = 8u << 20;  // bytes read per epoll callback before yielding
constexpr int kSocketBuffer = 4 << 20;

// Same-host single-copy path (see pair.h).
bool cmaEnabled() {
  static const bool on = [] {
    const char* v = std::getenv("GLB_TCP_CMA");
    return v == nullptr || std::atoi(v) != 0;
  }();
  return on;
}
size_t cmaMinBytes() {
  static const size_t n = [] {
    const char* v = std::getenv("GLB_TCP_CMA_MIN");
    long long b = v != nullptr ? std::atoll(v) : (256 << 10);
    return static_cast<size_t>(b < 1 ? 1 : b);
  }();
  return n;
}
std::atomic<uint64_t> g_cmaMessages{0};
std::atomic<uint64_t> g_cmaBytes{0};

// The word a peer must be able to read before it is allowed to pull from us.
uint64_t* probeWord() {
  static uint64_t word = [] {
    uint64_t v = 0;
    FILE* f = std::fopen("/dev/urandom", "rb");
    if (f != nullptr) {
      if (std::fread(&v, sizeof(v), 1, f) != 1) v = 0;
      std::fclose(f);
    }
    if (v == 0) v = 0x9e3779b97f4a7c15ull ^ (static_cast<uint64_t>(::getpid()) << 32) ^
                    static_cast<uint64_t>(std::chrono::steady_clock::now().time_since_epoch().count());
    return v;
  }();
  return &word;
}

// Not every mapping can be read with process_vm_readv: device-driver mappings (VM_IO /
// VM_PFNMAP, which is what pinned CUDA host allocations may be) make get_user_pages fail.
// The sender finds out by reading the first and last byte of the region from itself; if
// that does not work the payload goes through the socket as usual.
bool pullable(const char* data, size_t nbytes) {
  const pid_t self = ::getpid();  // not cached: a forked child must not read its parent
  char probe[2];
  struct iovec local[2] = {{&probe[0], 1}, {&probe[1], 1}};
  struct iovec remote[2] = {{const_cast<char*>(data), 1}, {const_cast<char*>(data) + nbytes - 1, 1}};
  return ::process_vm_readv(self, local, 2, remote, 2, 0) == 2;
}

bool retryableConnectError(int err) {
  return err == ECONNREFUSED || err == ETIMEDOUT || err == EHOSTUNREACH || err == ENETUNREACH ||
         err == ECONNRESET || err == EADDRNOTAVAIL || err == EINTR;
}
}  // namespace

Pair::Pair(Context* context, Device* device, int selfRank, int peerRank, std::chrono::milliseconds timeout,
           bool lazy)
    : context_(context),
      device_(device),
      selfRank_(selfRank),
      peerRank_(peerRank),
      timeout_(timeout),
      lazy_(lazy),
      loop_(&device->loop(device->nextLoopIndex())) {
  // Optionally derive the sequence number from (rank pair) so that addresses are
  // predictable across restarts (reference: GLOO_ENABLE_RANK_AS_SEQUENCE_NUMBER).
  self_ = device_->nextAddress();
}

void Pair::quiesce() {
  {
    std::unique_lock<std::mutex> lock(mu_);
    if (quiesced_) return;
    quiesced_ = true;
    if (expecting_) {
      device_->cancelExpectation(self_.seq());
      expecting_ = false;
    }
    if (state_ != CLOSED) {
      state_ = CLOSED;
      if (!failed_) {
        failed_ = true;
        exMsg_ = "pair destroyed";
      }
    }
    if (fd_ >= 0) loop_->removeDescriptor(fd_);
  }
  loop_->barrier();  // the loop may still be about to call handleEvents on us
  std::lock_guard<std::mutex> lock(mu_);  // ... or be inside it right now
}

Pair::~Pair() {
  quiesce();
  std::lock_guard<std::mutex> lock(mu_);
  if (fd_ >= 0) {
    ioShutdown();
    ::close(fd_);
    fd_ = -1;
  }
}

// ---- connection management --------------------------------------------------------

void Pair::connect(const std::vector<char>& bytes) {
  {
    std::lock_guard<std::mutex> g(mu_);
    GLB_ENFORCE(state_ == INITIALIZING, "connect() called twice on pair to rank ", peerRank_);
    peer_ = Address(bytes);
    havePeer_ = true;
    state_ = CONNECTING;
    if (selfRank_ < peerRank_) {
      // Listening side: the peer dials us and announces our sequence number.
      expecting_ = true;
    }
  }
  if (selfRank_ < peerRank_) {
    device_->expectConnection(self_.seq(), [this](Socket s) { attachSocket(std::move(s), false); });
  }
  if (!lazy_) ensureConnected(/*setup=*/true);
}

void Pair::ensureConnected(bool setup) {
  {
    std::unique_lock<std::mutex> lock(mu_);
    if (state_ == CONNECTED) return;
    if (setup && everConnected_ && closedByPeer_) return;
    throwIfException();
    GLB_ENFORCE(havePeer_, "pair to rank ", peerRank_, " has no peer address (connect() not called)");
    if (selfRank_ < peerRank_) {
      waitUntilConnected(lock, setup);
      return;
    }
  }
  dial();
}

void Pair::waitUntilConnected(std::unique_lock<std::mutex>& lock, bool setup) {
  auto pred = [&] { return state_ == CONNECTED || state_ == CLOSED; };
  if (timeout_ == kNoTimeout) {
    cv_.wait(lock, pred);
  } else {
    // Connecting can take a while when peers start at very different times;
    // the reference multiplies by 5 as well (tcp/pair.h:304-312).
    if (!cv_.wait_for(lock, timeout_ * 5, pred)) {
      signalException(strcat_all("Connect timeout waiting for rank ", peerRank_, " to dial ", self_.str()));
    }
  }
  if (setup && everConnected_ && closedByPeer_) return;
  throwIfException();
}

void Pair::dial() {
  // Serialise concurrent dial attempts from several user threads.
  std::lock_guard<std::mutex> dg(dialMu_);
  {
    std::lock_guard<std::mutex> g(mu_);
    if (state_ == CONNECTED) return;
    throwIfException();
  }
  const auto start = std::chrono::steady_clock::now();
  const auto& ss = peer_.sockaddr();
  const bool retries = !disableConnectionRetries();
  int attempt = 0;
  std::string lastErr;
  while (true) {
    attempt++;
    Socket sock = Socket::createForFamily(ss.ss_family);
    sock.setNonBlocking(true);
    int rv = ::connect(sock.fd(), reinterpret_cast<const struct sockaddr*>(&ss), sockaddrLen(ss));
    int err = rv == 0 ? 0 : errno;
    if (rv != 0 && err == EINPROGRESS) {
      struct pollfd pfd = {sock.fd(), POLLOUT, 0};
      int remainingMs = -1;
      if (timeout_ != kNoTimeout) {
        auto left = timeout_ - std::chrono::duration_cast<std::chrono::milliseconds>(
                                   std::chrono::steady_clock::now() - start);
        remainingMs = static_cast<int>(std::max<int64_t>(1, left.count()));
      }
      int prv;
      do {
        prv = ::poll(&pfd, 1, remainingMs);
      } while (prv == -1 && errno == EINTR);
      if (prv == 0) {
        err = ETIMEDOUT;
      } else {
        socklen_t len = sizeof(err);
        ::getsockopt(sock.fd(), SOL_SOCKET, SO_ERROR, &err, &len);
      }
    }
    if (err == 0) {
      attachSocket(std::move(sock), true);
      return;
    }
    lastErr = std::strerror(err);
    bool expired = timeout_ != kNoTimeout && std::chrono::steady_clock::now() - start > timeout_;
    if (!retries || !retryableConnectError(err) || expired) {
      std::lock_guard<std::mutex> g(mu_);
      signalException(strcat_all("connect to rank ", peerRank_, " at ", peer_.str(), " failed after ",
                                 attempt, " attempt(s): ", lastErr));
      throwIfException();
    }
    GLB_DEBUG("connect to ", peer_.str(), " failed (", lastErr, "), retrying");
    std::this_thread::sleep_for(std::chrono::milliseconds(std::min(attempt * 5, 100)));
  }
}

void Pair::attachSocket(Socket sock, bool initiator) {
  sock.setNonBlocking(true);
  sock.setNoDelay(true);
  sock.growBuffers(kSocketBuffer);
  if (initiator) {
    // Announce which of the listener's pairs this connection belongs to.
    Hello hello;
    hello.seq = peer_.seq();
    size_t off = 0;
    while (off < sizeof(hello)) {
      ssize_t n = ::send(sock.fd(), reinterpret_cast<const char*>(&hello) + off, sizeof(hello) - off, MSG_NOSIGNAL);
      if (n > 0) {
        off += static_cast<size_t>(n);
      } else if (n == -1 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
        struct pollfd pfd = {sock.fd(), POLLOUT, 0};
        ::poll(&pfd, 1, 100);
      } else if (n == -1 && errno == EINTR) {
        continue;
      } else {
        std::lock_guard<std::mutex> g(mu_);
        signalException(strcat_all("writing hello to rank ", peerRank_, ": ", std::strerror(errno)));
        return;
      }
    }
  }
  std::lock_guard<std::mutex> g(mu_);
  expecting_ = false;
  if (state_ == CLOSED) return;  // raced with close(); socket is dropped
  fd_ = sock.release();
  try {
    ioHandshake(initiator);
  } catch (const std::exception& e) {
    signalException(strcat_all("handshake with rank ", peerRank_, " failed: ", e.what()));
    cv_.notify_all();
    return;
  }
  state_ = CONNECTED;
  everConnected_ = true;
  if (!sync_) armEvents(false);
  if (cmaEnabled() && allowCma()) {
    TxOp caps;
    caps.hdr.opcode = OP_CAPS;
    caps.hdr.slot = static_cast<uint64_t>(::getpid());
    caps.hdr.aux = reinterpret_cast<uint64_t>(probeWord());
    caps.hdr.length = *probeWord();
    caps.bestEffort = true;
    try {
      enqueue(std::move(caps));
    } catch (const std::exception& e) {
      signalException(e.what());
    }
  }
  cv_.notify_all();
}

bool Pair::isConnected() {
  std::lock_guard<std::mutex> g(mu_);
  return state_ == CONNECTED;
}

void Pair::close() {
  int fd = -1;
  {
    std::unique_lock<std::mutex> g(mu_);
    if (expecting_) {
      device_->cancelExpectation(self_.seq());
      expecting_ = false;
    }
    if (state_ == CLOSED && fd_ < 0) return;
    if (fd_ >= 0) {
      // Abortive close: no TIME_WAIT, so test suites that churn through thousands
      // of connections do not exhaust ephemeral ports (reference: pair.cc:79-92).
      // An abortive close throws away what the kernel has not transmitted yet, and a
      // rank may legitimately close right after its last send completed (= was handed
      // to the kernel): give the queue a bounded moment to drain first.
      if (!failed_ && state_ == CONNECTED) {
        for (int i = 0; i < 2000 && !failed_ && state_ == CONNECTED && fd_ >= 0; i++) {
          int pending = 0;
          if (::ioctl(fd_, TIOCOUTQ, &pending) != 0 || pending == 0) break;
          g.unlock();  // the loop thread may still have to read from this pair meanwhile
          ::usleep(500);
          g.lock();
        }
      }
    }
    if (fd_ >= 0) {
      struct linger sl = {1, 0};
      ::setsockopt(fd_, SOL_SOCKET, SO_LINGER, &sl, sizeof(sl));
      loop_->removeDescriptor(fd_);
    }
    if (state_ != CLOSED) signalException("pair closed");
    fd = fd_;
  }
  loop_->barrier();
  std::lock_guard<std::mutex> g(mu_);
  if (fd >= 0 && fd_ == fd) {
    ioShutdown();
    ::close(fd_);
    fd_ = -1;
  }
}

void Pair::setSync(bool sync, bool busyPoll) {
  ensureConnected();
  std::lock_guard<std::mutex> g(mu_);
  throwIfException();
  if (sync == sync_) {
    busyPoll_ = busyPoll;
    return;
  }
  GLB_ENFORCE(sync, "cannot switch a pair back from sync to async mode");
  sync_ = true;
  busyPoll_ = busyPoll;
  // From here on the waiting thread drives the socket itself.
  loop_->removeDescriptor(fd_);
}

void Pair::armEvents(bool wantWrite) {
  wantWrite_ = wantWrite;
  loop_->modifyDescriptor(fd_, EPOLLIN | (wantWrite ? static_cast<int>(EPOLLOUT) : 0), this);
}

// ---- error handling ---------------------------------------------------------------

void Pair::throwIfException() {
  if (failed_) GLB_THROW_IO_EXCEPTION(exMsg_);
}

void Pair::signalExceptionExternal(const std::string& msg) {
  std::lock_guard<std::mutex> g(mu_);
  signalException(msg);
}

void Pair::signalException(const std::string& msg) {
  if (failed_) return;
  failed_ = true;
  exMsg_ = strcat_all("[rank ", selfRank_, " <-> rank ", peerRank_, "] ", msg);
  GLB_DEBUG("pair failure: ", exMsg_);
  state_ = CLOSED;
  if (fd_ >= 0) {
    if (!sync_) loop_->removeDescriptor(fd_);
    ::shutdown(fd_, SHUT_RDWR);  // unblocks the peer promptly; fd closed in close()/dtor
  }
  // Fail everything that is waiting on this pair.
  for (auto& op : tx_) {
    if (op.bbuf != nullptr) op.bbuf->signalException(exMsg_);
    if (op.hasUbuf && op.notify) {
      if (auto l = op.ubuf.lock()) l->signalException(exMsg_);
    }
  }
  tx_.clear();
  for (auto& op : awaitingFin_) {
    if (op.bbuf != nullptr) op.bbuf->signalException(exMsg_);
    if (op.hasUbuf && op.notify) {
      if (auto l = op.ubuf.lock()) l->signalException(exMsg_);
    }
  }
  awaitingFin_.clear();
  if (rx_.ubuf) rx_.ubuf->signalException(exMsg_);
  rx_.reset();
  for (auto& kv : recvBuffers_) kv.second->signalException(exMsg_);
  context_->failPostedRecvs(peerRank_, exMsg_);
  cv_.notify_all();
}

// ---- bound buffers ----------------------------------------------------------------

std::unique_ptr<::glb::transport::Buffer> Pair::createSendBuffer(int slot, void* ptr, size_t size) {
  return std::make_unique<Buffer>(this, slot, ptr, size, /*isRecv=*/false);
}

std::unique_ptr<::glb::transport::Buffer> Pair::createRecvBuffer(int slot, void* ptr, size_t size) {
  auto buf = std::make_unique<Buffer>(this, slot, ptr, size, /*isRecv=*/true);
  std::lock_guard<std::mutex> g(mu_);
  GLB_ENFORCE(recvBuffers_.find(slot) == recvBuffers_.end(), "recv buffer for slot ", slot, " already exists");
  recvBuffers_[slot] = buf.get();
  // Apply writes that raced ahead of the registration.
  auto it = parkedBound_.find(slot);
  if (it != parkedBound_.end()) {
    for (auto& pb : it->second) {
      GLB_ENFORCE_LE(pb.roffset + pb.data.size(), size, "parked write exceeds recv buffer");
      if (!pb.data.empty()) std::memcpy(static_cast<char*>(ptr) + pb.roffset, pb.data.data(), pb.data.size());
      buf->handleRecvCompletion();
    }
    parkedBound_.erase(it);
  }
  return buf;
}

void Pair::unregisterBuffer(Buffer* buf) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = recvBuffers_.find(buf->slot());
  if (it != recvBuffers_.end() && it->second == buf) recvBuffers_.erase(it);
  if (rx_.bbuf == buf) {
    // Destroyed mid-message: finish the message into a scratch area.
    rx_.stash.resize(rx_.hdr.nbytes);
    rx_.dst = rx_.stash.data();
    rx_.kind = RX_DISCARD;
    rx_.bbuf = nullptr;
  }
  bool poisoned = false;
  for (auto& op : tx_) {
    if (op.bbuf == buf) poisoned = true;
  }
  for (auto& op : awaitingFin_) {
    if (op.bbuf == buf) poisoned = true;
  }
  if (poisoned) signalException("bound buffer destroyed while a send was still queued");
}

void Pair::forgetUnbound(UnboundBuffer* buf) {
  std::lock_guard<std::mutex> g(mu_);
  bool poisoned = false;
  for (auto& op : tx_) {
    if (op.hasUbuf && op.ubufRaw == buf) poisoned = true;
  }
  for (auto& op : awaitingFin_) {
    if (op.hasUbuf && op.notify && op.ubufRaw == buf) poisoned = true;
  }
  if (poisoned) signalException("unbound buffer destroyed while a send was still queued");
}

// ---- send side --------------------------------------------------------------------

void Pair::sendBound(Buffer* buf, size_t offset, size_t length, size_t roffset) {
  ensureConnected();
  TxOp op;
  op.hdr.opcode = OP_SEND_BOUND;
  op.hdr.slot = static_cast<uint64_t>(buf->slot());
  op.hdr.nbytes = length;
  op.hdr.roffset = roffset;
  op.bbuf = buf;
  std::lock_guard<std::mutex> g(mu_);
  setPayload(op, static_cast<const char*>(buf->ptr()) + offset, length);
  enqueue(std::move(op));
}

void Pair::sendUnbound(UnboundBuffer* buf, uint64_t slot, size_t offset, size_t nbytes) {
  TxOp op;
  op.hdr.opcode = OP_SEND_UNBOUND;
  op.hdr.slot = slot;
  op.hdr.nbytes = nbytes;
  op.hasUbuf = true;
  op.ubuf = buf->weak();
  op.ubufRaw = buf;
  std::lock_guard<std::mutex> g(mu_);
  setPayload(op, static_cast<const char*>(buf->ptr) + offset, nbytes);
  enqueue(std::move(op));
}

void Pair::sendPut(UnboundBuffer* buf, uint64_t regionId, size_t offset, size_t roffset, size_t nbytes) {
  TxOp op;
  op.hdr.opcode = OP_PUT;
  op.hdr.aux = regionId;
  op.hdr.nbytes = nbytes;
  op.hdr.roffset = roffset;
  op.hasUbuf = true;
  op.ubuf = buf->weak();
  op.ubufRaw = buf;
  std::lock_guard<std::mutex> g(mu_);
  setPayload(op, static_cast<const char*>(buf->ptr) + offset, nbytes);
  enqueue(std::move(op));
}

void Pair::sendGetRequest(uint64_t requestId, uint64_t regionId, size_t roffset, size_t nbytes) {
  TxOp op;
  op.hdr.opcode = OP_GET_REQ;
  op.hdr.slot = requestId;
  op.hdr.aux = regionId;
  op.hdr.roffset = roffset;
  op.hdr.length = nbytes;
  std::lock_guard<std::mutex> g(mu_);
  enqueue(std::move(op));
}

void Pair::send(::glb::transport::UnboundBuffer* tbuf, uint64_t tag, size_t offset, size_t nbytes) {
  auto* buf = dynamic_cast<UnboundBuffer*>(tbuf);
  GLB_ENFORCE(buf != nullptr, "not a tcp unbound buffer");
  ensureConnected();
  sendUnbound(buf, tag, offset, nbytes);
}

void Pair::recv(::glb::transport::UnboundBuffer* tbuf, uint64_t tag, size_t offset, size_t nbytes) {
  auto* buf = dynamic_cast<UnboundBuffer*>(tbuf);
  GLB_ENFORCE(buf != nullptr, "not a tcp unbound buffer");
  context_->postRecv(buf, {peerRank_}, tag, offset, nbytes);
}

void Pair::setPayload(TxOp& op, const char* data, size_t nbytes) {
  op.hdr.nbytes = nbytes;
  op.data = data;
  if (peerCanPull_ && nbytes >= cmaMinBytes() && pullable(data, nbytes)) {
    // Header only; the receiver pulls the bytes and answers FIN.
    op.cma = true;
    op.cmaId = ++cmaSeq_;
    op.hdr.flags |= F_CMA;
    op.hdr.length = reinterpret_cast<uint64_t>(data);
    // The id travels in whichever header field the opcode leaves free.
    if (op.hdr.opcode == OP_PUT) {
      op.hdr.slot = op.cmaId;
    } else {
      op.hdr.aux = op.cmaId;
    }
    op.nbytes = 0;
  } else {
    op.nbytes = nbytes;
  }
}

void Pair::wroteTx(TxOp&& op) {
  if (op.cma) {
    awaitingFin_.push_back(std::move(op));
  } else {
    completeTx(op);
  }
}

void Pair::enqueue(TxOp&& op) {
  throwIfException();
  GLB_ENFORCE(state_ == CONNECTED, "pair to rank ", peerRank_, " is not connected");
  if (sync_) {
    // Blocking write from the calling thread.
    while (!tryWrite(op)) {
      throwIfException();
      struct pollfd pfd = {fd_, POLLOUT, 0};
      if (!busyPoll_) ::poll(&pfd, 1, 100);
    }
    wroteTx(std::move(op));
    return;
  }
  if (tx_.empty() && tryWrite(op)) {
    wroteTx(std::move(op));
    return;
  }
  throwIfException();
  tx_.push_back(std::move(op));
  if (!wantWrite_) armEvents(true);
}

bool Pair::tryWrite(TxOp& op) {
  const size_t total = sizeof(WireHeader) + op.nbytes;
  Lease<UnboundBuffer> lease;
  if (op.hasUbuf && op.nbytes > 0) {
    lease = op.ubuf.lock();
    if (!lease) {
      signalException("unbound buffer destroyed while a send was in flight");
      return false;
    }
  }
  while (op.sent < total) {
    struct iovec iov[2];
    int cnt = 0;
    if (op.sent < sizeof(WireHeader)) {
      iov[cnt].iov_base = reinterpret_cast<char*>(&op.hdr) + op.sent;
      iov[cnt].iov_len = sizeof(WireHeader) - op.sent;
      cnt++;
      if (op.nbytes > 0) {
        iov[cnt].iov_base = const_cast<char*>(op.data);
        iov[cnt].iov_len = op.nbytes;
        cnt++;
      }
    } else {
      size_t poff = op.sent - sizeof(WireHeader);
      iov[cnt].iov_base = const_cast<char*>(op.data) + poff;
      iov[cnt].iov_len = op.nbytes - poff;
      cnt++;
    }
    ssize_t n = ioSend(iov, cnt);
    if (n >= 0) {
      op.sent += static_cast<size_t>(n);
      continue;
    }
    if (errno == EINTR) continue;
    if (errno == EAGAIN || errno == EWOULDBLOCK) return false;
    if (op.bestEffort) {
      // The peer finished and closed before this unsolicited frame went out. Whatever it
      // sent is still in our receive queue and must stay deliverable; if more was expected
      // from it, the read side reports the closed connection.
      op.sent = total;
      return true;
    }
    signalException(strcat_all("send: ", std::strerror(errno), " (peer ", peer_.str(), ")"));
    return false;
  }
  return true;
}

void Pair::completeTx(TxOp& op) {
  if (op.bbuf != nullptr) op.bbuf->handleSendCompletion();
  if (op.hasUbuf && op.notify) {
    if (auto l = op.ubuf.lock()) l->handleSendCompletion(peerRank_);
  }
}

void Pair::flushTx() {
  while (!tx_.empty()) {
    if (!tryWrite(tx_.front())) return;  // EAGAIN (or failure, which cleared tx_)
    TxOp op = std::move(tx_.front());
    tx_.pop_front();
    wroteTx(std::move(op));
  }
  if (wantWrite_ && state_ == CONNECTED) armEvents(false);
}

// ---- receive side -----------------------------------------------------------------

ssize_t Pair::ioRecv(void* buf, size_t len) { return ::recv(fd_, buf, len, MSG_DONTWAIT); }

ssize_t Pair::ioSend(const struct iovec* iov, int iovcnt) {
  struct msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = const_cast<struct iovec*>(iov);
  msg.msg_iovlen = static_cast<size_t>(iovcnt);
  return ::sendmsg(fd_, &msg, MSG_NOSIGNAL | MSG_DONTWAIT);
}

void Pair::handleEvents(int events) {
  std::lock_guard<std::mutex> g(mu_);
  if (state_ != CONNECTED || sync_) return;
  try {
    if (events & EPOLLOUT) flushTx();
    if (state_ == CONNECTED && (events & (EPOLLIN | EPOLLHUP | EPOLLERR | EPOLLRDHUP))) readLoop(kReadBudget);
  } catch (const std::exception& e) {
    signalException(e.what());
  }
}

void Pair::readLoop(size_t budget) {
  size_t consumed = 0;
  while (state_ == CONNECTED && (consumed < budget || ioPending())) {
    if (rx_.hdrRead < sizeof(WireHeader)) {
      ssize_t n = ioRecv(reinterpret_cast<char*>(&rx_.hdr) + rx_.hdrRead, sizeof(WireHeader) - rx_.hdrRead);
      if (n > 0) {
        rx_.hdrRead += static_cast<size_t>(n);
        consumed += static_cast<size_t>(n);
        if (rx_.hdrRead < sizeof(WireHeader)) continue;
        if (rx_.hdr.magic != WireHeader::kMagic) {
          signalException("protocol error: bad message magic");
          return;
        }
        beginMessage();
        if (state_ != CONNECTED) return;
      } else if (n == 0) {
        closedByPeer_ = rx_.hdrRead == 0;
        signalException(strcat_all("Connection closed by peer [", peer_.str(), "]"));
        return;
      } else if (errno == EINTR) {
        continue;
      } else if (errno == EAGAIN || errno == EWOULDBLOCK) {
        return;
      } else {
        closedByPeer_ = errno == ECONNRESET && rx_.hdrRead == 0;
        signalException(strcat_all("recv: ", std::strerror(errno), " (peer ", peer_.str(), ")"));
        return;
      }
    }
    if ((rx_.hdr.flags & F_CMA) != 0 && rx_.payloadRead < rx_.hdr.nbytes) {
      if (rx_.deferred) {
        rx_.payloadRead = rx_.hdr.nbytes;  // parked: the posting thread pulls it later
      } else {
        if (!pullPayload()) return;
        consumed += rx_.hdr.nbytes;
      }
    }
    if (rx_.payloadRead < rx_.hdr.nbytes) {
      ssize_t n = ioRecv(rx_.dst + rx_.payloadRead, rx_.hdr.nbytes - rx_.payloadRead);
      if (n > 0) {
        rx_.payloadRead += static_cast<size_t>(n);
        consumed += static_cast<size_t>(n);
      } else if (n == 0) {
        signalException(strcat_all("Connection closed by peer [", peer_.str(), "] mid-message"));
        return;
      } else if (errno == EINTR) {
        continue;
      } else if (errno == EAGAIN || errno == EWOULDBLOCK) {
        return;
      } else {
        signalException(strcat_all("recv: ", std::strerror(errno), " (peer ", peer_.str(), ")"));
        return;
      }
    }
    if (rx_.payloadRead == rx_.hdr.nbytes) {
      const bool fin = (rx_.hdr.flags & F_CMA) != 0 && !rx_.deferred;
      const uint64_t finId = rx_.hdr.opcode == OP_PUT ? rx_.hdr.slot : rx_.hdr.aux;
      finishMessage();
      rx_.reset();
      if (fin && state_ == CONNECTED) {
        TxOp op;
        op.hdr.opcode = OP_FIN;
        op.hdr.slot = finId;
        enqueue(std::move(op));
      }
    }
  }
}

bool Pair::copyFrom(char* dst, uint64_t srcAddr, size_t nbytes, std::string* err) {
  const char* src = reinterpret_cast<const char*>(srcAddr);
  size_t done = 0;
  while (done < nbytes) {
    struct iovec local = {dst + done, nbytes - done};
    struct iovec remote = {const_cast<char*>(src) + done, nbytes - done};
    ssize_t n = ::process_vm_readv(peerPid_, &local, 1, &remote, 1, 0);
    if (n > 0) {
      done += static_cast<size_t>(n);
    } else if (n == -1 && errno == EINTR) {
      continue;
    } else {
      *err = strcat_all("process_vm_readv from rank ", peerRank_, " (pid ", peerPid_, "): ",
                        n == 0 ? "short read" : std::strerror(errno));
      return false;
    }
  }
  g_cmaMessages.fetch_add(1, std::memory_order_relaxed);
  g_cmaBytes.fetch_add(nbytes, std::memory_order_relaxed);
  return true;
}

bool Pair::pullPayload() {
  if (!canPull_) {
    signalException("protocol error: peer used the single-copy path without permission");
    return false;
  }
  std::string err;
  if (!copyFrom(rx_.dst + rx_.payloadRead, rx_.hdr.length + rx_.payloadRead, rx_.hdr.nbytes - rx_.payloadRead, &err)) {
    signalException(err);
    return false;
  }
  rx_.payloadRead = rx_.hdr.nbytes;
  return true;
}

void Pair::pullDeferred(char* dst, uint64_t srcAddr, size_t nbytes, uint64_t id) {
  // peerPid_ / canPull_ were fixed by the CAPS exchange long before any single-copy
  // message could exist; the copy itself touches no pair state, so it runs unlocked and
  // the loop thread keeps serving this pair meanwhile.
  std::string err;
  const bool ok = copyFrom(dst, srcAddr, nbytes, &err);
  std::lock_guard<std::mutex> g(mu_);
  if (!ok) {
    signalException(err);
    throwIfException();
  }
  throwIfException();
  TxOp op;
  op.hdr.opcode = OP_FIN;
  op.hdr.slot = id;
  enqueue(std::move(op));
}

uint64_t Pair::cmaMessages() { return g_cmaMessages.load(std::memory_order_relaxed); }
uint64_t Pair::cmaBytes() { return g_cmaBytes.load(std::memory_order_relaxed); }

void Pair::beginMessage() {
  const auto& h = rx_.hdr;
  switch (h.opcode) {
    case OP_SEND_UNBOUND: {
      Context::Match m;
      bool matched;
      if ((h.flags & F_CMA) != 0) {
        if (!canPull_) {
          signalException("protocol error: peer used the single-copy path without permission");
          return;
        }
        Context::RemotePayload rp;
        rp.srcAddr = h.length;
        rp.nbytes = h.nbytes;
        rp.id = h.aux;
        matched = context_->matchOrDefer(peerRank_, h.slot, rp, &m);
        if (!matched) {
          rx_.kind = RX_NONE;
          rx_.deferred = true;
          break;
        }
      } else {
        matched = context_->matchIncoming(peerRank_, h.slot, &m);
      }
      if (matched) {
        if (h.nbytes > m.capacity) {
          m.buf->signalException(strcat_all("distributed collective mismatch: rank ", peerRank_, " sent ",
                                            h.nbytes, " bytes on slot ", h.slot, " but the posted recv holds ",
                                            m.capacity));
          rx_.stash.resize(h.nbytes);
          rx_.dst = rx_.stash.data();
          rx_.kind = RX_DISCARD;
          break;
        }
        rx_.kind = RX_UNBOUND_DIRECT;
        rx_.dst = m.dst;
        rx_.ubuf = std::move(m.buf);
      } else {
        rx_.kind = RX_UNBOUND_UNEXPECTED;
        rx_.stash.resize(h.nbytes);
        rx_.dst = rx_.stash.data();
      }
      break;
    }
    case OP_SEND_BOUND: {
      auto it = recvBuffers_.find(static_cast<int>(h.slot));
      if (it != recvBuffers_.end()) {
        Buffer* b = it->second;
        if (!rangeFits(h.roffset, h.nbytes, b->size())) {
          signalException(strcat_all("bound write out of range on slot ", h.slot, ": offset ", h.roffset,
                                     " + ", h.nbytes, " > ", b->size()));
          return;
        }
        rx_.kind = RX_BOUND_DIRECT;
        rx_.bbuf = b;
        rx_.dst = static_cast<char*>(b->ptr()) + h.roffset;
      } else {
        rx_.kind = RX_BOUND_UNEXPECTED;
        rx_.stash.resize(h.nbytes);
        rx_.dst = rx_.stash.data();
      }
      break;
    }
    case OP_PUT: {
      Lease<UnboundBuffer> lease;
      if (context_->lookupRegion(h.aux, &lease) && rangeFits(h.roffset, h.nbytes, lease->size)) {
        rx_.kind = RX_PUT;
        rx_.dst = static_cast<char*>(lease->ptr) + h.roffset;
        rx_.ubuf = std::move(lease);
      } else {
        GLB_WARN("dropping put to unknown or too small region ", h.aux, " from rank ", peerRank_);
        rx_.kind = RX_DISCARD;
        rx_.stash.resize(h.nbytes);
        rx_.dst = rx_.stash.data();
      }
      break;
    }
    case OP_GET_REQ: {
      Lease<UnboundBuffer> lease;
      TxOp op;
      op.hdr.opcode = OP_GET_RESP;
      op.hdr.slot = h.slot;
      if (context_->lookupRegion(h.aux, &lease) && rangeFits(h.roffset, h.length, lease->size)) {
        setPayload(op, static_cast<const char*>(lease->ptr) + h.roffset, h.length);
        op.hasUbuf = true;
        op.notify = false;
        op.ubuf = lease->weak();
        op.ubufRaw = lease.get();
      } else {
        GLB_WARN("get from unknown or too small region ", h.aux, " by rank ", peerRank_);
        op.hdr.flags = F_ERROR;  // no payload
      }
      lease.release();
      rx_.kind = RX_NONE;
      enqueue(std::move(op));
      break;
    }
    case OP_GET_RESP: {
      Context::Match m;
      if (context_->takePendingGet(h.slot, &m)) {
        if ((h.flags & F_ERROR) != 0 || h.nbytes != m.capacity) {
          m.buf->signalException("one-sided get failed on the remote side (bad key or range)");
          rx_.kind = RX_DISCARD;
          rx_.stash.resize(h.nbytes);
          rx_.dst = rx_.stash.data();
        } else {
          rx_.kind = RX_GET_RESP;
          rx_.dst = m.dst;
          rx_.ubuf = std::move(m.buf);
        }
      } else {
        rx_.kind = RX_DISCARD;
        rx_.stash.resize(h.nbytes);
        rx_.dst = rx_.stash.data();
      }
      break;
    }
    case OP_CAPS: {
      rx_.kind = RX_NONE;
      if (!cmaEnabled() || !allowCma()) break;
      // The pid and probe address come from the peer: only a peer on THIS host (same address at
      // both ends of the connection, or loopback) may name a process to read from.
      if (!peerIsThisHost(fd_)) break;
      peerPid_ = static_cast<int>(h.slot);
      uint64_t seen = 0;
      struct iovec local = {&seen, sizeof(seen)};
      struct iovec remote = {reinterpret_cast<void*>(h.aux), sizeof(seen)};
      ssize_t n = ::process_vm_readv(peerPid_, &local, 1, &remote, 1, 0);
      if (n == static_cast<ssize_t>(sizeof(seen)) && seen == h.length) {
        canPull_ = true;
        TxOp op;
        op.hdr.opcode = OP_CAPS_OK;
        op.bestEffort = true;
        enqueue(std::move(op));
      } else {
        GLB_DEBUG("single-copy path to rank ", peerRank_, " unavailable (different host or no ptrace permission)");
      }
      break;
    }
    case OP_CAPS_OK:
      rx_.kind = RX_NONE;
      peerCanPull_ = cmaEnabled() && allowCma();
      break;
    case OP_FIN: {
      rx_.kind = RX_NONE;
      if (awaitingFin_.empty()) {
        signalException("protocol error: FIN without a pending single-copy send");
        return;
      }
      auto it = awaitingFin_.begin();
      while (it != awaitingFin_.end() && it->cmaId != h.slot) ++it;
      if (it == awaitingFin_.end()) {
        signalException("protocol error: FIN for an unknown single-copy send");
        return;
      }
      TxOp op = std::move(*it);
      awaitingFin_.erase(it);
      completeTx(op);
      break;
    }
    default:
      signalException(strcat_all("protocol error: unknown opcode ", h.opcode));
      return;
  }
}

void Pair::finishMessage() {
  switch (rx_.kind) {
    case RX_UNBOUND_DIRECT:
    case RX_GET_RESP:
      rx_.ubuf->handleRecvCompletion(peerRank_);
      break;
    case RX_UNBOUND_UNEXPECTED:
      context_->deliverUnexpected(peerRank_, rx_.hdr.slot, std::move(rx_.stash));
      break;
    case RX_BOUND_DIRECT:
      rx_.bbuf->handleRecvCompletion();
      break;
    case RX_BOUND_UNEXPECTED: {
      // The recv buffer may have been registered while the payload was in flight.
      auto it = recvBuffers_.find(static_cast<int>(rx_.hdr.slot));
      if (it != recvBuffers_.end()) {
        Buffer* b = it->second;
        if (rx_.hdr.roffset + rx_.stash.size() > b->size()) {
          signalException("bound write out of range");
          return;
        }
        if (!rx_.stash.empty()) {
          std::memcpy(static_cast<char*>(b->ptr()) + rx_.hdr.roffset, rx_.stash.data(), rx_.stash.size());
        }
        b->handleRecvCompletion();
      } else {
        parkedBound_[static_cast<int>(rx_.hdr.slot)].push_back(ParkedBound{rx_.hdr.roffset, std::move(rx_.stash)});
      }
      break;
    }
    case RX_PUT:
    case RX_DISCARD:
    case RX_NONE:
      break;
  }
}

// ---- sync mode --------------------------------------------------------------------

void Pair::syncWait(std::unique_lock<std::mutex>& lock, const std::function<bool()>& pred,
                    std::chrono::milliseconds timeout, const char* what) {
  const auto start = std::chrono::steady_clock::now();
  while (true) {
    if (failed_ || pred()) return;
    if (state_ != CONNECTED) return;
    if (!busyPoll_) {
      struct pollfd pfd = {fd_, POLLIN, 0};
      lock.unlock();
      ::poll(&pfd, 1, 20);
      lock.lock();
      if (failed_ || pred()) return;
    }
    try {
      readLoop(kReadBudget);
    } catch (const std::exception& e) {
      signalException(e.what());
      return;
    }
    if (timeout != kNoTimeout && std::chrono::steady_clock::now() - start > timeout) {
      signalException(strcat_all("Timed out waiting ", timeout.count(), "ms for ", what, " operation to complete"));
      return;
    }
  }
}

int64_t Pair::spinBudgetNanos() {
  static const int64_t ns = [] {
    const char* v = std::getenv("GLB_TCP_SPIN_US");
    long us = v != nullptr ? std::strtol(v, nullptr, 10) : 1000;
    if (us < 0) us = 0;
    return static_cast<int64_t>(us) * 1000;
  }();
  return ns;
}

namespace {
inline void cpuRelax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield" ::: "memory");
#endif
}
}  // namespace

void Pair::spinWait(std::unique_lock<std::mutex>& lock, const std::function<bool()>& pred) {
  const int64_t budget = spinBudgetNanos();
  if (budget == 0 || sync_ || failed_ || state_ != CONNECTED || pred()) return;
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::nanoseconds(budget);
  unsigned spins = 0;
  // While this thread polls the socket the loop thread would only wake up to find the
  // data already taken (level-triggered epoll fires on every arrival): park the
  // descriptor for the duration of the spin.
  PauseGuard quiet(this, /*locked=*/true);
  while (!failed_ && state_ == CONNECTED && !pred()) {
    try {
      readLoop(kReadBudget);
    } catch (const std::exception& e) {
      signalException(e.what());
      return;
    }
    if (failed_ || pred()) return;
    // Let the loop thread / senders in, then look again. On a saturated machine the
    // thread that has to produce what we are waiting for may be runnable but not
    // running: give the core away now and then instead of burning the whole time slice.
    lock.unlock();
    if ((++spins & 31) == 0) {
      ::sched_yield();
    } else {
      for (int i = 0; i < 16; i++) cpuRelax();
    }
    lock.lock();
    if (std::chrono::steady_clock::now() >= deadline) return;
  }
}

// ---- epoll parking while a waiter polls --------------------------------------------------

bool Pair::pauseEventsLocked() {
  if (sync_ || wantWrite_ || state_ != CONNECTED || fd_ < 0) return false;
  if (paused_++ == 0) loop_->modifyDescriptor(fd_, 0, this);
  return true;
}

void Pair::resumeEventsLocked() {
  // armEvents() is idempotent: if a queued write re-enabled the descriptor meanwhile this
  // just registers the same interest again.
  if (paused_ > 0 && --paused_ == 0 && !sync_ && state_ == CONNECTED && fd_ >= 0) armEvents(wantWrite_);
}

Pair::PauseGuard::PauseGuard(Pair* p, bool locked) : pair_(p), locked_(locked) {
  static const bool enabled = [] {
    const char* v = std::getenv("GLB_TCP_PARK_EPOLL");
    return v == nullptr || std::atoi(v) != 0;
  }();
  if (!enabled) return;
  if (locked_) {
    took_ = pair_->pauseEventsLocked();
  } else {
    std::lock_guard<std::mutex> g(pair_->mu_);
    took_ = pair_->pauseEventsLocked();
  }
}

Pair::PauseGuard::~PauseGuard() {
  if (!took_) return;
  if (locked_) {
    pair_->resumeEventsLocked();
  } else {
    std::lock_guard<std::mutex> g(pair_->mu_);
    pair_->resumeEventsLocked();
  }
}

void Pair::tryProgress() {
  std::unique_lock<std::mutex> lock(mu_, std::try_to_lock);
  if (!lock.owns_lock() || failed_ || state_ != CONNECTED) return;
  try {
    readLoop(kReadBudget);
  } catch (const std::exception& e) {
    signalException(e.what());
  }
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
