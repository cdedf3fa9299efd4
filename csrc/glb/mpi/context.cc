#include "glb/mpi/context.h"

#if GLB_USE_MPI

#include <cstring>
#include <mutex>
#include <vector>

#include "glb/common/logging.h"
#include "glb/transport/context.h"

namespace glb {
namespace mpi {

static int rankOf(const MPI_Comm& c) {
  int r;
  GLB_ENFORCE_EQ(MPI_Comm_rank(c, &r), MPI_SUCCESS);
  return r;
}
static int sizeOf(const MPI_Comm& c) {
  int s;
  GLB_ENFORCE_EQ(MPI_Comm_size(c, &s), MPI_SUCCESS);
  return s;
}

MPIScope::MPIScope() {
  int inited = 0;
  MPI_Initialized(&inited);
  if (!inited) {
    int provided = 0;
    GLB_ENFORCE_EQ(MPI_Init_thread(nullptr, nullptr, MPI_THREAD_MULTIPLE, &provided), MPI_SUCCESS);
  }
}
MPIScope::~MPIScope() {
  int fin = 0;
  MPI_Finalized(&fin);
  if (!fin) MPI_Finalize();
}

std::shared_ptr<Context> Context::createManaged() {
  static std::mutex mu;
  static std::weak_ptr<MPIScope> weak;
  std::shared_ptr<MPIScope> scope;
  {
    std::lock_guard<std::mutex> g(mu);
    scope = weak.lock();
    if (!scope) {
      scope = std::make_shared<MPIScope>();
      weak = scope;
    }
  }
  auto ctx = std::make_shared<Context>(MPI_COMM_WORLD);
  ctx->scope_ = scope;
  return ctx;
}

Context::Context(const MPI_Comm& comm) : ::glb::Context(rankOf(comm), sizeOf(comm)) {
  GLB_ENFORCE_EQ(MPI_Comm_dup(comm, &comm_), MPI_SUCCESS);
}

Context::~Context() { MPI_Comm_free(&comm_); }

// Every rank exports one rendezvous blob; sizes via MPI_Allreduce(MAX), blobs via
// MPI_Allgather of fixed-size slots, then the transport context connects from them.
void Context::connectFullMesh(std::shared_ptr<transport::Device>& dev) {
  auto tctx = dev->createContext(rank, size);
  tctx->setTimeout(getTimeout());
  auto blob = tctx->exportRendezvousBlob();
  int mine = static_cast<int>(blob.size()) + static_cast<int>(sizeof(int));
  int maxLen = 0;
  GLB_ENFORCE_EQ(MPI_Allreduce(&mine, &maxLen, 1, MPI_INT, MPI_MAX, comm_), MPI_SUCCESS);
  std::vector<char> slot(maxLen, 0), all(static_cast<size_t>(maxLen) * size);
  int len = static_cast<int>(blob.size());
  std::memcpy(slot.data(), &len, sizeof(int));
  std::memcpy(slot.data() + sizeof(int), blob.data(), blob.size());
  GLB_ENFORCE_EQ(MPI_Allgather(slot.data(), maxLen, MPI_BYTE, all.data(), maxLen, MPI_BYTE, comm_), MPI_SUCCESS);
  std::vector<std::vector<char>> blobs(size);
  for (int i = 0; i < size; i++) {
    int l = 0;
    std::memcpy(&l, all.data() + static_cast<size_t>(i) * maxLen, sizeof(int));
    blobs[i].assign(all.data() + static_cast<size_t>(i) * maxLen + sizeof(int),
                    all.data() + static_cast<size_t>(i) * maxLen + sizeof(int) + l);
  }
  tctx->connectWithBlobs(blobs);
  device_ = dev;
  transportContext_ = std::move(tctx);
}

}  // namespace mpi
}  // namespace glb

#endif  // GLB_USE_MPI
