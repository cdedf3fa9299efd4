#include "mpi.h"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <vector>

namespace {
int gSize = 1;
thread_local int tRank = 0;
std::atomic<int> gInit{0}, gFinal{0}, gLive{0};

// Reusable rendezvous of all rank threads: everyone deposits, the last arrival publishes.
struct Exchange {
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<std::vector<char>> slots;
  std::vector<char> result;

  std::vector<char> run(const void* data, size_t n, bool maxInt) {
    std::unique_lock<std::mutex> lock(mu);
    if (slots.size() != static_cast<size_t>(gSize)) slots.assign(gSize, {});
    slots[tRank].assign(static_cast<const char*>(data), static_cast<const char*>(data) + n);
    const long gen = generation;
    if (++arrived == gSize) {
      result.clear();
      if (maxInt) {
        int best = 0;
        for (int r = 0; r < gSize; r++) {
          int v;
          std::memcpy(&v, slots[r].data(), sizeof(v));
          if (r == 0 || v > best) best = v;
        }
        result.resize(sizeof(int));
        std::memcpy(result.data(), &best, sizeof(int));
      } else {
        for (int r = 0; r < gSize; r++) result.insert(result.end(), slots[r].begin(), slots[r].end());
      }
      arrived = 0;
      generation++;
      cv.notify_all();
      return result;
    }
    cv.wait(lock, [&] { return generation != gen; });
    return result;
  }
} gEx;
}  // namespace

extern "C" {
void fake_mpi_world(int size) { gSize = size; }
void fake_mpi_set_rank(int rank) { tRank = rank; }
int fake_mpi_live_comms(void) { return gLive.load(); }

int MPI_Initialized(int* flag) { *flag = gInit.load() > 0; return MPI_SUCCESS; }
int MPI_Finalized(int* flag) { *flag = gFinal.load() > 0; return MPI_SUCCESS; }
int MPI_Init_thread(int*, char***, int required, int* provided) {
  gInit++;
  *provided = required;
  return MPI_SUCCESS;
}
int MPI_Finalize(void) { gFinal++; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = tRank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = gSize; return MPI_SUCCESS; }
int MPI_Comm_dup(MPI_Comm, MPI_Comm* out) { *out = 100 + gLive.fetch_add(1); return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm* comm) { gLive--; *comm = -1; return MPI_SUCCESS; }
int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm) {
  if (count != 1 || dt != MPI_INT || op != MPI_MAX) return 1;
  auto r = gEx.run(send, sizeof(int), true);
  std::memcpy(recv, r.data(), sizeof(int));
  return MPI_SUCCESS;
}
int MPI_Allgather(const void* send, int scount, MPI_Datatype sdt, void* recv, int rcount, MPI_Datatype rdt, MPI_Comm) {
  if (sdt != MPI_BYTE || rdt != MPI_BYTE || scount != rcount) return 1;
  auto r = gEx.run(send, static_cast<size_t>(scount), false);
  std::memcpy(recv, r.data(), r.size());
  return MPI_SUCCESS;
}
}
