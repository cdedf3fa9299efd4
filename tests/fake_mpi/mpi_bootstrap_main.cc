// Runs glb::mpi::Context end to end against the thread-world MPI in this directory:
// N rank threads build the full mesh with MPI_Allreduce(MAX) + MPI_Allgather instead of a
// store, then allreduce over the resulting TCP mesh. Prints "OK <sum>" on success.
#include <cstdio>
#include <thread>
#include <vector>

#include "glb/allreduce.h"
#include "glb/barrier.h"
#include "glb/math.h"
#include "glb/mpi/context.h"
#include "glb/transport/tcp/device.h"
#include "mpi.h"

int main(int argc, char** argv) {
  const int P = argc > 1 ? std::atoi(argv[1]) : 3;
  fake_mpi_world(P);
  std::vector<float> results(P, 0.f);
  std::vector<std::thread> ths;
  for (int r = 0; r < P; r++) {
    ths.emplace_back([&, r] {
      fake_mpi_set_rank(r);
      auto ctx = r == 0 ? glb::mpi::Context::createManaged() : std::make_shared<glb::mpi::Context>(MPI_COMM_WORLD);
      auto dev = glb::transport::tcp::CreateDevice(glb::transport::tcp::attr("127.0.0.1"));
      ctx->connectFullMesh(dev);
      std::vector<float> x(1000, static_cast<float>(r + 1));
      glb::AllreduceOptions opts(ctx);
      opts.setOutput(x.data(), x.size());
      opts.setReduceFunction([](void* c, const void* a, const void* b, size_t n) { glb::sum<float>(c, a, b, n); });
      glb::allreduce(opts);
      results[r] = x[999];
      glb::BarrierOptions b(ctx);
      glb::barrier(b);
    });
  }
  for (auto& t : ths) t.join();
  const float want = P * (P + 1) / 2.0f;
  for (int r = 0; r < P; r++) {
    if (results[r] != want) {
      std::printf("FAIL rank %d got %f want %f\n", r, results[r], want);
      return 1;
    }
  }
  if (fake_mpi_live_comms() != 0) {
    std::printf("FAIL %d communicators leaked\n", fake_mpi_live_comms());
    return 1;
  }
  int fin = 0;
  MPI_Finalized(&fin);
  std::printf("OK %g finalized=%d\n", want, fin);
  return 0;
}
