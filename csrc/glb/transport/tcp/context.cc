#include "glb/transport/tcp/context.h"

#include <cstring>

#include "glb/common/logging.h"
#include "glb/common/utils.h"
#include "glb/transport/tcp/pair.h"
#include "glb/transport/tcp/unbound_buffer.h"

namespace glb {
namespace transport {
namespace tcp {

std::string RemoteKey::serialize() const { return strcat_all(rank, ",", size, ",", regionId); }

std::unique_ptr<RemoteKey> RemoteKey::deserialize(const std::string& s) {
  int rank = 0;
  unsigned long long size = 0, id = 0;
  GLB_ENFORCE(std::sscanf(s.c_str(), "%d,%llu,%llu", &rank, &size, &id) == 3, "malformed remote key: ", s);
  return std::make_unique<RemoteKey>(rank, static_cast<size_t>(size), static_cast<uint64_t>(id));
}

Context::Context(std::shared_ptr<Device> device, int rank, int size)
    : ::glb::transport::Context(rank, size), device_(std::move(device)), unexpected_(size) {}

Context::~Context() {
  // Pairs reference this context; make sure they are gone (and detached from the
  // loop) before the matching tables are destroyed.
  for (auto& p : pairs_) p.reset();
}

std::unique_ptr<::glb::transport::Pair>& Context::createPair(int peer) {
  GLB_ENFORCE(peer >= 0 && peer < size && peer != rank, "invalid peer rank ", peer);
  pairs_[peer] = std::make_unique<Pair>(this, device_.get(), rank, peer, getTimeout(), device_->isLazy());
  return pairs_[peer];
}

std::unique_ptr<::glb::transport::Pair>& Context::getPair(int peer) {
  auto& p = pairs_.at(peer);
  if (p && device_->isLazy()) static_cast<Pair*>(p.get())->ensureConnected();
  return p;
}

Pair* Context::tcpPair(int peer) {
  GLB_ENFORCE(peer >= 0 && peer < size, "rank out of range: ", peer);
  GLB_ENFORCE_NE(peer, rank, "no pair to self");
  auto& p = pairs_[peer];
  GLB_ENFORCE(p != nullptr, "pair to rank ", peer, " was not created");
  auto* tp = static_cast<Pair*>(p.get());
  tp->ensureConnected();
  return tp;
}

std::unique_ptr<::glb::transport::UnboundBuffer> Context::createUnboundBuffer(void* ptr, size_t size) {
  return std::make_unique<UnboundBuffer>(shared_from_this(), ptr, size);
}

// Blob: [u32 ver][u32 size][u32 hostLen][host][u32 addrLen][sockaddr][u64 seq x size]
// seq[i] is the sequence number of OUR pair that talks to rank i.
std::vector<char> Context::makeBlob() {
  std::vector<char> blob;
  auto put = [&](const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    blob.insert(blob.end(), c, c + n);
  };
  auto put32 = [&](uint32_t v) { put(&v, 4); };
  const std::string host = getHostname();
  put32(1);
  put32(static_cast<uint32_t>(size));
  put32(static_cast<uint32_t>(host.size()));
  put(host.data(), host.size());
  const auto& ss = device_->sockaddr();
  put32(static_cast<uint32_t>(sizeof(ss)));
  put(&ss, sizeof(ss));
  for (int i = 0; i < size; i++) {
    uint64_t seq = 0;
    if (i != rank) {
      if (!pairs_[i]) createPair(i);
      seq = static_cast<Pair*>(pairs_[i].get())->address().seq();
    }
    put(&seq, 8);
  }
  return blob;
}

void Context::connectToPeerBlob(int peer, const std::vector<char>& blob, const std::string& selfHost,
                                int* localRank) {
  size_t off = 0;
  auto get = [&](void* p, size_t n) {
    GLB_ENFORCE_LE(off + n, blob.size(), "truncated rendezvous blob from rank ", peer);
    std::memcpy(p, blob.data() + off, n);
    off += n;
  };
  uint32_t ver, n, hostLen, addrLen;
  get(&ver, 4);
  GLB_ENFORCE_EQ(ver, 1u, "rendezvous blob version mismatch");
  get(&n, 4);
  GLB_ENFORCE_EQ(static_cast<int>(n), size, "rank ", peer, " disagrees on world size");
  get(&hostLen, 4);
  std::string host(hostLen, '\0');
  get(&host[0], hostLen);
  get(&addrLen, 4);
  struct sockaddr_storage ss;
  GLB_ENFORCE_EQ(addrLen, sizeof(ss));
  get(&ss, sizeof(ss));
  std::vector<uint64_t> seqs(size);
  get(seqs.data(), 8 * size);
  if (peer < rank && host == selfHost) (*localRank)++;
  Address remote(ss, seqs[rank]);
  pairs_[peer]->connect(remote.bytes());
}

void Context::createAndConnectAllPairs(std::shared_ptr<IStore> store) {
  auto blob = makeBlob();
  store->set(std::to_string(rank), blob);

  const std::string host = getHostname();
  int localRank = 0;
  std::vector<int> peers;
  for (int i = 0; i < size; i++) {
    if (i != rank) peers.push_back(i);
  }
  const bool batched = store->has_extended_api() && isStoreExtendedApiEnabled();
  constexpr size_t kBatch = 128;
  for (size_t b = 0; b < peers.size(); b += kBatch) {
    size_t e = std::min(peers.size(), b + kBatch);
    std::vector<std::vector<char>> blobs;
    if (batched) {
      std::vector<std::string> keys;
      for (size_t i = b; i < e; i++) keys.push_back(std::to_string(peers[i]));
      store->wait(keys, getTimeout());
      blobs = store->multi_get(keys);
    } else {
      for (size_t i = b; i < e; i++) blobs.push_back(store->wait_get(std::to_string(peers[i]), getTimeout()));
    }
    for (size_t i = b; i < e; i++) connectToPeerBlob(peers[i], blobs[i - b], host, &localRank);
  }
  for (int i = 0; i < size; i++) {
    if (pairs_[i]) pairs_[i]->setLocalRank(localRank);
  }
  if (logLevel() >= LogLevel::INFO) {
    GLB_INFO("rank ", rank, "/", size, " connected to ", peers.size(), " peers via ", device_->str(),
             device_->isLazy() ? " (lazy)" : "", ", localRank=", localRank);
  }
}

std::vector<char> Context::exportRendezvousBlob() { return makeBlob(); }

void Context::connectWithBlobs(const std::vector<std::vector<char>>& blobs) {
  GLB_ENFORCE_EQ(static_cast<int>(blobs.size()), size);
  const std::string host = getHostname();
  int localRank = 0;
  for (int i = 0; i < size; i++) {
    if (i == rank) continue;
    if (!pairs_[i]) createPair(i);
    connectToPeerBlob(i, blobs[i], host, &localRank);
  }
  for (int i = 0; i < size; i++) {
    if (pairs_[i]) pairs_[i]->setLocalRank(localRank);
  }
}

void Context::signalException(const std::string& msg) {
  for (int i = 0; i < size; i++) {
    if (pairs_[i]) static_cast<Pair*>(pairs_[i].get())->signalExceptionExternal(msg);
  }
}

// ---- matching ------------------------------------------------------------------

void Context::postRecv(UnboundBuffer* buf, std::vector<int> srcRanks, uint64_t slot, size_t offset,
                       size_t nbytes) {
  GLB_ENFORCE(!srcRanks.empty(), "recv needs at least one source rank");
  for (int r : srcRanks) {
    GLB_ENFORCE(r >= 0 && r < size && r != rank, "invalid source rank ", r);
  }
  // A message that already arrived is delivered even if its pair has since seen the
  // peer close (a rank that finished its part of a collective may exit while slower
  // ranks still have its data parked here). Only when nothing is waiting do the
  // connections have to be alive — and, in lazy mode, created.
  bool waiting = false;
  auto somethingWaiting = [&] {
    std::lock_guard<std::mutex> g(matchMu_);
    for (int r : srcRanks) {
      auto it = unexpected_[r].find(slot);
      if (it != unexpected_[r].end() && !it->second.empty()) return true;
    }
    return false;
  };
  waiting = somethingWaiting();
  if (!waiting) {
    try {
      for (int r : srcRanks) tcpPair(r);
    } catch (const IoException&) {
      // The liveness check can block behind the read that delivers the very message
      // this recv is for, followed by the peer's EOF. Nothing more can arrive after
      // that, so a second look at the queue is conclusive.
      if (!somethingWaiting()) throw;
    }
  }
  int matchedRank = -1;
  bool deferred = false;
  RemotePayload pull;
  {
    std::lock_guard<std::mutex> g(matchMu_);
    // An already-arrived message wins; rotate the scan start for fairness.
    const size_t n = srcRanks.size();
    const size_t start = n > 1 ? static_cast<size_t>(anyCursor_++ % n) : 0;
    for (size_t k = 0; k < n && matchedRank < 0; k++) {
      int r = srcRanks[(start + k) % n];
      auto it = unexpected_[r].find(slot);
      if (it == unexpected_[r].end() || it->second.empty()) continue;
      auto& u = it->second.front();
      GLB_ENFORCE_LE(u.size(), nbytes, "distributed collective mismatch: rank ", r,
                     " sent more bytes on slot ", slot, " than the posted recv holds");
      if (u.deferred) {
        pull = u.remote;
        deferred = true;
      } else if (!u.data.empty()) {
        std::memcpy(static_cast<char*>(buf->ptr) + offset, u.data.data(), u.data.size());
      }
      it->second.pop_front();
      if (it->second.empty()) unexpected_[r].erase(it);
      matchedRank = r;
    }
    if (matchedRank < 0) {
      posted_[slot].push_back(PostedRecv{buf, buf->weak(), offset, nbytes, std::move(srcRanks)});
      return;
    }
  }
  if (deferred) {
    // The sender's bytes are still in its address space: fetch them straight into the
    // destination from this (the posting) thread, then release the sender.
    auto* pair = static_cast<Pair*>(pairs_[matchedRank].get());
    pair->pullDeferred(static_cast<char*>(buf->ptr) + offset, pull.srcAddr, pull.nbytes, pull.id);
  }
  buf->handleRecvCompletion(matchedRank);
}

bool Context::matchOrDefer(int srcRank, uint64_t slot, const RemotePayload& payload, Match* out) {
  std::lock_guard<std::mutex> g(matchMu_);
  auto& uq = unexpected_[srcRank][slot];
  if (uq.empty()) {
    auto it = posted_.find(slot);
    if (it != posted_.end()) {
      auto& q = it->second;
      for (auto p = q.begin(); p != q.end();) {
        if (!p->accepts(srcRank)) {
          ++p;
          continue;
        }
        auto lease = p->buf.lock();
        if (!lease) {
          p = q.erase(p);
          continue;
        }
        out->dst = static_cast<char*>(lease->ptr) + p->offset;
        out->capacity = p->nbytes;
        out->buf = std::move(lease);
        q.erase(p);
        if (q.empty()) posted_.erase(it);
        if (uq.empty()) unexpected_[srcRank].erase(slot);
        return true;
      }
    }
  }
  Unexpected u;
  u.deferred = true;
  u.remote = payload;
  unexpected_[srcRank][slot].push_back(std::move(u));
  return false;
}

bool Context::matchIncoming(int srcRank, uint64_t slot, Match* out) {
  std::lock_guard<std::mutex> g(matchMu_);
  // FIFO per (src, slot): if older unexpected messages from this source are still
  // queued on this slot the new one must queue behind them.
  auto uit = unexpected_[srcRank].find(slot);
  if (uit != unexpected_[srcRank].end() && !uit->second.empty()) return false;
  auto it = posted_.find(slot);
  if (it == posted_.end()) return false;
  auto& q = it->second;
  for (auto p = q.begin(); p != q.end();) {
    if (!p->accepts(srcRank)) {
      ++p;
      continue;
    }
    auto lease = p->buf.lock();
    if (!lease) {  // buffer died with the recv still posted
      p = q.erase(p);
      continue;
    }
    out->dst = static_cast<char*>(lease->ptr) + p->offset;
    out->capacity = p->nbytes;
    out->buf = std::move(lease);
    q.erase(p);
    if (q.empty()) posted_.erase(it);
    return true;
  }
  return false;
}

void Context::deliverUnexpected(int srcRank, uint64_t slot, std::vector<char>&& data) {
  Lease<UnboundBuffer> lease;
  {
    std::lock_guard<std::mutex> g(matchMu_);
    auto& uq = unexpected_[srcRank][slot];
    if (uq.empty()) {
      auto it = posted_.find(slot);
      if (it != posted_.end()) {
        auto& q = it->second;
        for (auto p = q.begin(); p != q.end(); ++p) {
          if (!p->accepts(srcRank)) continue;
          auto l = p->buf.lock();
          if (!l) continue;
          if (data.size() > p->nbytes) {
            l->signalException(strcat_all("distributed collective mismatch: rank ", srcRank, " sent ",
                                          data.size(), " bytes, recv posted for ", p->nbytes));
          } else {
            if (!data.empty()) std::memcpy(static_cast<char*>(l->ptr) + p->offset, data.data(), data.size());
            lease = std::move(l);
          }
          q.erase(p);
          if (q.empty()) posted_.erase(it);
          break;
        }
      }
    }
    if (!lease) {
      Unexpected u;
      u.data = std::move(data);
      uq.push_back(std::move(u));
      return;
    }
    if (uq.empty()) unexpected_[srcRank].erase(slot);
  }
  lease->handleRecvCompletion(srcRank);
}

void Context::cancelPostedRecvs(UnboundBuffer* buf) {
  std::lock_guard<std::mutex> g(matchMu_);
  for (auto it = posted_.begin(); it != posted_.end();) {
    auto& q = it->second;
    for (auto p = q.begin(); p != q.end();) {
      p = (p->raw == buf) ? q.erase(p) : p + 1;
    }
    it = q.empty() ? posted_.erase(it) : std::next(it);
  }
}

void Context::failPostedRecvs(int srcRank, const std::string& msg) {
  std::vector<Lease<UnboundBuffer>> victims;
  {
    std::lock_guard<std::mutex> g(matchMu_);
    for (auto it = posted_.begin(); it != posted_.end();) {
      auto& q = it->second;
      for (auto p = q.begin(); p != q.end();) {
        if (p->accepts(srcRank)) {
          auto l = p->buf.lock();
          if (l) victims.push_back(std::move(l));
          p = q.erase(p);
        } else {
          ++p;
        }
      }
      it = q.empty() ? posted_.erase(it) : std::next(it);
    }
  }
  for (auto& v : victims) v->signalException(msg);
}

// ---- one-sided -----------------------------------------------------------------

uint64_t Context::registerRegion(UnboundBuffer* buf) {
  uint64_t id = nextId_.fetch_add(1);
  std::lock_guard<std::mutex> g(regionMu_);
  regions_.emplace(id, std::make_pair(buf, buf->weak()));
  return id;
}

bool Context::lookupRegion(uint64_t id, Lease<UnboundBuffer>* lease) {
  std::lock_guard<std::mutex> g(regionMu_);
  auto it = regions_.find(id);
  if (it == regions_.end()) return false;
  *lease = it->second.second.lock();
  return static_cast<bool>(*lease);
}

uint64_t Context::registerPendingGet(UnboundBuffer* buf, size_t offset, size_t nbytes) {
  uint64_t id = nextId_.fetch_add(1);
  std::lock_guard<std::mutex> g(regionMu_);
  pendingGets_.emplace(id, PendingGet{buf, buf->weak(), offset, nbytes});
  return id;
}

bool Context::takePendingGet(uint64_t id, Match* out) {
  std::lock_guard<std::mutex> g(regionMu_);
  auto it = pendingGets_.find(id);
  if (it == pendingGets_.end()) return false;
  auto lease = it->second.buf.lock();
  bool ok = static_cast<bool>(lease);
  if (ok) {
    out->dst = static_cast<char*>(lease->ptr) + it->second.offset;
    out->capacity = it->second.nbytes;
    out->buf = std::move(lease);
  }
  pendingGets_.erase(it);
  return ok;
}

void Context::forgetBuffer(UnboundBuffer* buf, uint64_t regionId) {
  cancelPostedRecvs(buf);
  {
    std::lock_guard<std::mutex> g(regionMu_);
    if (regionId != 0) regions_.erase(regionId);
    for (auto it = pendingGets_.begin(); it != pendingGets_.end();) {
      it = (it->second.raw == buf) ? pendingGets_.erase(it) : std::next(it);
    }
  }
  for (int i = 0; i < size; i++) {
    if (pairs_[i]) static_cast<Pair*>(pairs_[i].get())->forgetUnbound(buf);
  }
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
