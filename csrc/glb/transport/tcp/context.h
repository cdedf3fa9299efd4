// tcp::Context — pairs, lean rendezvous (one blob per rank), receiver-side message
// matching (posted receives vs. unexpected arrivals, recv-from-any), the region
// table behind the software one-sided put/get, and error fan-out.
// Parity: gloo/transport/tcp/context.{h,cc} + the Tally machinery of
// gloo/transport/context.h:111-290 (replaced by local matching).
#pragma once

#include <atomic>
#include <deque>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "glb/common/memory.h"
#include "glb/transport/context.h"
#include "glb/transport/tcp/device.h"

namespace glb {
namespace transport {
namespace tcp {

class Pair;
class UnboundBuffer;

class RemoteKey : public ::glb::transport::RemoteKey {
 public:
  RemoteKey(int rank, size_t size, uint64_t regionId)
      : ::glb::transport::RemoteKey(rank, size), regionId(regionId) {}
  std::string serialize() const override;
  static std::unique_ptr<RemoteKey> deserialize(const std::string& s);
  const uint64_t regionId;
};

class Context : public ::glb::transport::Context, public std::enable_shared_from_this<Context> {
 public:
  Context(std::shared_ptr<Device> device, int rank, int size);
  ~Context() override;

  std::unique_ptr<::glb::transport::Pair>& getPair(int rank) override;
  std::unique_ptr<::glb::transport::Pair>& createPair(int rank) override;
  void createAndConnectAllPairs(std::shared_ptr<IStore> store) override;
  std::unique_ptr<::glb::transport::UnboundBuffer> createUnboundBuffer(void* ptr, size_t size) override;
  std::unique_ptr<::glb::transport::RemoteKey> deserializeRemoteKey(const std::string& s) override {
    return RemoteKey::deserialize(s);
  }
  std::vector<char> exportRendezvousBlob() override;
  void connectWithBlobs(const std::vector<std::vector<char>>& blobs) override;

  Pair* tcpPair(int rank);  // connected pair or throws

  // Poison every pair (and all pending operations) of this context.
  void signalException(const std::string& msg);

  // ---- matching (called by UnboundBuffer and Pair) ---------------------------
  struct Match {
    Lease<UnboundBuffer> buf;
    char* dst = nullptr;
    size_t capacity = 0;
  };
  void postRecv(UnboundBuffer* buf, std::vector<int> srcRanks, uint64_t slot, size_t offset, size_t nbytes);
  // Loop thread, header just arrived: returns a match if a recv is posted.
  bool matchIncoming(int srcRank, uint64_t slot, Match* out);
  // Loop thread, unexpected payload fully read.
  void deliverUnexpected(int srcRank, uint64_t slot, std::vector<char>&& data);
  // Single-copy message (header only on the wire): match it like matchIncoming, or park
  // its descriptor so that the recv, once posted, pulls the bytes straight into place.
  struct RemotePayload {
    uint64_t srcAddr = 0;
    size_t nbytes = 0;
    uint64_t id = 0;  // echoed in the FIN that completes the sender's operation
  };
  bool matchOrDefer(int srcRank, uint64_t slot, const RemotePayload& payload, Match* out);
  void cancelPostedRecvs(UnboundBuffer* buf);
  void failPostedRecvs(int srcRank, const std::string& msg);

  // ---- one-sided -------------------------------------------------------------
  uint64_t registerRegion(UnboundBuffer* buf);
  bool lookupRegion(uint64_t id, Lease<UnboundBuffer>* lease);
  uint64_t registerPendingGet(UnboundBuffer* buf, size_t offset, size_t nbytes);
  bool takePendingGet(uint64_t id, Match* out);
  void forgetBuffer(UnboundBuffer* buf, uint64_t regionId);

 private:
  struct PostedRecv {
    UnboundBuffer* raw;
    WeakAnchor<UnboundBuffer> buf;
    size_t offset;
    size_t nbytes;
    std::vector<int> srcRanks;
    bool accepts(int r) const {
      for (int s : srcRanks) {
        if (s == r) return true;
      }
      return false;
    }
  };
  struct PendingGet {
    UnboundBuffer* raw;
    WeakAnchor<UnboundBuffer> buf;
    size_t offset;
    size_t nbytes;
  };

  std::vector<char> makeBlob();
  void connectToPeerBlob(int peer, const std::vector<char>& blob, const std::string& selfHost, int* localRank);

  std::shared_ptr<Device> device_;
  std::vector<sequence_number_t> expectedSeq_;

  std::mutex matchMu_;
  std::unordered_map<uint64_t, std::deque<PostedRecv>> posted_;
  // unexpected_[src][slot] -> FIFO of messages that arrived before their recv: either the
  // payload itself or, for single-copy messages, where to pull it from.
  struct Unexpected {
    std::vector<char> data;
    bool deferred = false;
    RemotePayload remote;
    size_t size() const { return deferred ? remote.nbytes : data.size(); }
  };
  std::vector<std::unordered_map<uint64_t, std::deque<Unexpected>>> unexpected_;
  uint64_t anyCursor_ = 0;  // rotates the scan start of recv-from-any for fairness

  std::mutex regionMu_;
  std::unordered_map<uint64_t, std::pair<UnboundBuffer*, WeakAnchor<UnboundBuffer>>> regions_;
  std::unordered_map<uint64_t, PendingGet> pendingGets_;
  std::atomic<uint64_t> nextId_{1};
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
