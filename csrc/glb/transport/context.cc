#include "glb/transport/context.h"

#include <cstring>

#include "glb/common/logging.h"
#include "glb/common/utils.h"

namespace glb {
namespace transport {

// Blob layout: [u32 naddr] then naddr x ([u32 len][bytes]) — address i is the one
// of the pair that talks to rank i (empty for self).
void Context::createAndConnectAllPairs(std::shared_ptr<IStore> store) {
  const std::string host = getHostname();
  store->set(strcat_all("host_", rank), toBytes(host));

  std::vector<char> blob;
  auto put32 = [&](uint32_t v) {
    const char* p = reinterpret_cast<const char*>(&v);
    blob.insert(blob.end(), p, p + 4);
  };
  put32(static_cast<uint32_t>(size));
  for (int i = 0; i < size; i++) {
    if (i == rank) {
      put32(0);
      continue;
    }
    auto& pair = createPair(i);
    auto bytes = pair->address().bytes();
    put32(static_cast<uint32_t>(bytes.size()));
    blob.insert(blob.end(), bytes.begin(), bytes.end());
  }
  store->set(strcat_all("addr_", rank), blob);

  int localRank = 0;
  for (int i = 0; i < size; i++) {
    if (i == rank) continue;
    auto peerHost = toString(store->wait_get(strcat_all("host_", i), getTimeout()));
    if (i < rank && peerHost == host) localRank++;
    auto peer = store->wait_get(strcat_all("addr_", i), getTimeout());
    size_t off = 0;
    auto get32 = [&]() {
      GLB_ENFORCE_LE(off + 4, peer.size(), "truncated rendezvous blob");
      uint32_t v;
      std::memcpy(&v, peer.data() + off, 4);
      off += 4;
      return v;
    };
    uint32_t n = get32();
    GLB_ENFORCE_EQ(static_cast<int>(n), size, "peer ", i, " has a different world size");
    std::vector<char> mine;
    for (uint32_t j = 0; j < n; j++) {
      uint32_t len = get32();
      if (static_cast<int>(j) == rank) mine.assign(peer.begin() + off, peer.begin() + off + len);
      off += len;
    }
    getPair(i)->connect(mine);
  }
  for (int i = 0; i < size; i++) {
    if (i != rank) getPair(i)->setLocalRank(localRank);
  }
}

}  // namespace transport
}  // namespace glb
