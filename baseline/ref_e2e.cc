// End-to-end step of the UNMODIFIED reference library (pytorch/gloo), through its own public
// API and stock code path: every step copies the rank's input from pinned host memory to
// the GPU, runs gloo::CudaAllreduceRingChunked<float>::run() (default workspace =
// CudaHostWorkspace, exactly what the reference's benchmark_cuda runs), and copies the
// result back to pinned host memory. Nothing of gloo_b200 is linked here: the program is
// built by baseline/build_reference.sh against the reference's own static libraries.
//
//   ref_e2e --size N --rank R --shared-path DIR --elements E --steps K --warmup W
// prints one line:  REF_E2E rank=R ms_per_step=<mean> steps=K
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "gloo/barrier_all_to_all.h"
#include "gloo/cuda_allreduce_ring_chunked.h"
#include "gloo/rendezvous/context.h"
#include "gloo/rendezvous/file_store.h"
#include "gloo/transport/tcp/device.h"

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));           \
      return 2;                                                               \
    }                                                                         \
  } while (0)

int main(int argc, char** argv) {
  int size = 1, rank = 0, steps = 5, warmup = 2;
  long elements = 1000;
  std::string path = "/tmp/ref_e2e";
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    const char* v = argv[i + 1];
    if (k == "--size") size = std::atoi(v);
    else if (k == "--rank") rank = std::atoi(v);
    else if (k == "--shared-path") path = v;
    else if (k == "--elements") elements = std::atol(v);
    else if (k == "--steps") steps = std::atoi(v);
    else if (k == "--warmup") warmup = std::atoi(v);
  }
  CK(cudaSetDevice(0));  // the launcher narrows CUDA_VISIBLE_DEVICES to this rank's GPU
  float *hin = nullptr, *hout = nullptr, *dbuf = nullptr;
  const size_t bytes = static_cast<size_t>(elements) * sizeof(float);
  CK(cudaMallocHost(reinterpret_cast<void**>(&hin), bytes));
  CK(cudaMallocHost(reinterpret_cast<void**>(&hout), bytes));
  CK(cudaMalloc(reinterpret_cast<void**>(&dbuf), bytes));
  for (long i = 0; i < elements; i++) hin[i] = 1.0f;
  cudaStream_t stream;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

  gloo::transport::tcp::attr attr;
  attr.iface = "lo";
  auto dev = gloo::transport::tcp::CreateDevice(attr);
  auto store = std::make_shared<gloo::rendezvous::FileStore>(path);
  auto ctx = std::make_shared<gloo::rendezvous::Context>(rank, size);
  ctx->connectFullMesh(store, dev);

  std::vector<float*> ptrs{dbuf};
  std::vector<cudaStream_t> streams{stream};
  gloo::CudaAllreduceRingChunked<float> algo(ctx, ptrs, static_cast<int>(elements), streams);
  gloo::BarrierAllToAll barrier(ctx);

  auto step = [&]() -> int {
    CK(cudaMemcpyAsync(dbuf, hin, bytes, cudaMemcpyHostToDevice, stream));
    algo.run();
    CK(cudaMemcpyAsync(hout, dbuf, bytes, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  };
  for (int i = 0; i < warmup; i++) {
    if (step()) return 2;
  }
  if (size > 1) barrier.run();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < steps; i++) {
    if (step()) return 2;
  }
  const auto t1 = std::chrono::steady_clock::now();
  if (size > 1) barrier.run();
  const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count() / steps;
  if (hout[0] != static_cast<float>(size) || hout[elements - 1] != static_cast<float>(size)) {
    std::fprintf(stderr, "ref_e2e: wrong result %f (expected %d)\n", hout[0], size);
    return 3;
  }
  std::printf("REF_E2E rank=%d ms_per_step=%.5f steps=%d h2d_bytes=%zu d2h_bytes=%zu\n", rank, ms, steps, bytes, bytes);
  return 0;
}
