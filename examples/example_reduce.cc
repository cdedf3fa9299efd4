// C++ API example (counterpart of gloo/examples/example_reduce.cc): file rendezvous,
// TCP mesh, new-style reduce to rank 0 with a custom reduction function.
//   g++ -std=c++17 -Icsrc examples/example_reduce.cc -Lgloo_b200/lib -lglb -o example_reduce
//   ./example_reduce <rank> <size> <dir>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "glb/reduce.h"
#include "glb/rendezvous/context.h"
#include "glb/rendezvous/file_store.h"
#include "glb/transport/tcp/device.h"

int main(int argc, char** argv) {
  if (argc != 4) return 1;
  const int rank = std::atoi(argv[1]), size = std::atoi(argv[2]);
  auto dev = glb::transport::tcp::CreateDevice("127.0.0.1");
  auto store = std::make_shared<glb::rendezvous::FileStore>(argv[3]);
  auto ctx = std::make_shared<glb::rendezvous::Context>(rank, size);
  ctx->connectFullMesh(store, dev);

  std::vector<int> in(4, rank + 1), out(4, 0);
  glb::ReduceOptions opts(ctx);
  opts.setInput(in.data(), in.size());
  opts.setOutput(out.data(), out.size());
  opts.setRoot(0);
  opts.setReduceFunction([](void* c, const void* a, const void* b, size_t n) {
    for (size_t i = 0; i < n; i++) static_cast<int*>(c)[i] = static_cast<const int*>(a)[i] + static_cast<const int*>(b)[i];
  });
  glb::reduce(opts);
  if (rank == 0) std::printf("sum = %d\n", out[0]);
  ctx->closeConnections();
  return 0;
}
