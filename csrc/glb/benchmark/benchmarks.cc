// Benchmark registry. Host names match gloo/benchmark/main.cc:1036-1041, CUDA names
// match gloo/benchmark/cuda_main.cc:173-214; the cuda_allgather / cuda_alltoall(_v) /
// cuda_reduce_scatter / cuda_reduce / cuda_barrier entries are net-new (the
// reference has no CUDA variant of those collectives).
#include <cuda_runtime.h>

#include <cstring>
#include <numeric>

#include "glb/allgather.h"
#include "glb/allgather_ring.h"
#include "glb/allgatherv.h"
#include "glb/allreduce.h"
#include "glb/allreduce_bcube.h"
#include "glb/allreduce_halving_doubling.h"
#include "glb/allreduce_local.h"
#include "glb/allreduce_ring.h"
#include "glb/allreduce_ring_chunked.h"
#include "glb/alltoall.h"
#include "glb/alltoallv.h"
#include "glb/barrier.h"
#include "glb/barrier_all_to_all.h"
#include "glb/benchmark/harness.h"
#include "glb/broadcast.h"
#include "glb/broadcast_one_to_all.h"
#include "glb/cuda/algorithms.h"
#include "glb/cuda/collectives.h"
#include "glb/cuda/kernels.h"
#include "glb/gather.h"
#include "glb/pairwise_exchange.h"
#include "glb/reduce.h"
#include "glb/reduce_scatter.h"
#include "glb/reduce_scatter_halving_doubling.h"
#include "glb/scatter.h"
#include "glb/transport/unbound_buffer.h"

namespace glb {
namespace benchmark {

namespace {

// Input pattern of the reference (benchmark.h:52-70): memory[j] = j*(P*inputs) + rank*inputs + i,
// so an allreduce yields j*stride^2 + stride(stride-1)/2 with stride = P*inputs.
template <typename T>
struct HostData {
  std::vector<std::vector<T>> bufs;
  std::vector<T*> ptrs;
  void fill(int rank, int size, int inputs, size_t n) {
    const size_t stride = static_cast<size_t>(size) * inputs;
    bufs.assign(inputs, std::vector<T>(std::max<size_t>(n, 1)));
    ptrs.clear();
    for (int i = 0; i < inputs; i++) {
      for (size_t j = 0; j < n; j++) bufs[i][j] = T(static_cast<float>(j * stride + rank * inputs + i));
      ptrs.push_back(bufs[i].data());
    }
  }
};

template <typename T>
void checkAllreduce(const std::vector<T*>& ptrs, size_t n, int size, int inputs) {
  const double stride = static_cast<double>(size) * inputs;
  const size_t limit = std::min<size_t>(n, 1 << 16);
  for (auto* p : ptrs) {
    for (size_t j = 0; j < limit; j++) {
      const double exp = j * stride * stride + stride * (stride - 1) / 2;
      if (sizeof(T) == 2 && exp > 60000.0) break;  // beyond the range of a 16-bit float: nothing to compare
      const double got = static_cast<double>(static_cast<float>(p[j]));
      const double tol = sizeof(T) == 2 ? 1e-2 * std::max(1.0, exp) : 1e-5 * std::max(1.0, exp);
      GLB_ENFORCE(std::abs(got - exp) <= tol, "Mismatch at index ", j, ": got ", got, " expected ", exp);
    }
  }
}

template <typename T, typename Algo>
Benchmark hostAllreduce(std::shared_ptr<Context> ctx, const Options& o) {
  auto data = std::make_shared<HostData<T>>();
  auto algo = std::make_shared<std::unique_ptr<Algorithm>>();
  auto count = std::make_shared<size_t>(0);
  Benchmark b;
  b.elementSize = sizeof(T);
  b.busFactor = ctx->size > 1 ? 2.0 * (ctx->size - 1) / ctx->size : 0.0;
  b.initialize = [=](size_t n) {
    *count = n;
    data->fill(ctx->rank, ctx->size, o.inputs, n);
    algo->reset(new Algo(ctx, data->ptrs, n));
  };
  b.run = [=] { (*algo)->run(); };
  b.verify = [=] { checkAllreduce<T>(data->ptrs, *count, ctx->size, o.inputs); };
  return b;
}

template <typename T>
Benchmark newAllreduce(std::shared_ptr<Context> ctx, const Options& o, AllreduceOptions::Algorithm alg) {
  auto data = std::make_shared<HostData<T>>();
  auto outs = std::make_shared<HostData<T>>();
  auto count = std::make_shared<size_t>(0);
  Benchmark b;
  b.elementSize = sizeof(T);
  b.busFactor = ctx->size > 1 ? 2.0 * (ctx->size - 1) / ctx->size : 0.0;
  b.initialize = [=](size_t n) {
    *count = n;
    data->fill(ctx->rank, ctx->size, o.inputs, n);
    outs->fill(0, 1, o.inputs, n);
  };
  b.run = [=] {
    AllreduceOptions opts(ctx);
    opts.setAlgorithm(alg);
    opts.setInputs(data->ptrs, *count);
    opts.setOutputs(outs->ptrs, *count);
    opts.setReduceFunction([](void* c, const void* a, const void* bb, size_t n) { sum<T>(c, a, bb, n); });
    allreduce(opts);
  };
  b.verify = [=] { checkAllreduce<T>(outs->ptrs, *count, ctx->size, o.inputs); };
  return b;
}

template <typename T, typename F>
Benchmark simple(std::shared_ptr<Context> ctx, double busFactor, size_t inMul, size_t outMul, F body) {
  auto in = std::make_shared<std::vector<T>>();
  auto out = std::make_shared<std::vector<T>>();
  auto count = std::make_shared<size_t>(0);
  Benchmark b;
  b.elementSize = sizeof(T);
  b.busFactor = busFactor;
  b.initialize = [=](size_t n) {
    *count = n;
    in->assign(std::max<size_t>(1, n * inMul), T(static_cast<float>(ctx->rank)));
    out->assign(std::max<size_t>(1, n * outMul), T(0.0f));
  };
  b.run = [=] { body(ctx, in->data(), out->data(), *count); };
  return b;
}

// ---- CUDA ---------------------------------------------------------------------------------

int pickDevice(const std::shared_ptr<Context>& ctx, const Options& o) {
  int n = cuda::deviceCount();
  GLB_ENFORCE_GT(n, 0, "no CUDA device visible");
  return o.cudaDevice >= 0 ? o.cudaDevice : ctx->rank % n;
}

cuda::AllreduceAlgo parseCudaAlgo(const std::string& s, cuda::AllreduceAlgo dflt) {
  if (s == "auto") return dflt;
  if (s == "one_shot") return cuda::AllreduceAlgo::ONE_SHOT;
  if (s == "two_shot") return cuda::AllreduceAlgo::TWO_SHOT;
  if (s == "nvls") return cuda::AllreduceAlgo::NVLS;
  if (s == "literal") return dflt;
  GLB_THROW(Exception, "unknown --cuda-algo ", s);
}

struct CudaBuffers {
  std::vector<void*> ptrs;
  int device = 0;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ~CudaBuffers() {
    cudaSetDevice(device);
    for (auto* p : ptrs) cudaFree(p);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (stream) cudaStreamDestroy(stream);
  }
  void alloc(int dev, int inputs, size_t nbytes) {
    device = dev;
    bytes = nbytes;
    GLB_CUDA_CHECK(cudaSetDevice(dev));
    for (int i = 0; i < inputs; i++) {
      void* p = nullptr;
      GLB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(nbytes, 16)));
      ptrs.push_back(p);
    }
    GLB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    GLB_CUDA_CHECK(cudaEventCreate(&e0));
    GLB_CUDA_CHECK(cudaEventCreate(&e1));
  }
};

template <typename T>
Benchmark cudaAllreduce(std::shared_ptr<Context> ctx, const Options& o, cuda::AllreduceAlgo named) {
  auto bufs = std::make_shared<CudaBuffers>();
  auto core = std::make_shared<std::unique_ptr<cuda::CudaAllreduceCore>>();
  auto count = std::make_shared<size_t>(0);
  const DataType dt = DataTypeOf<T>::value;
  Benchmark b;
  b.elementSize = sizeof(T);
  b.busFactor = ctx->size > 1 ? 2.0 * (ctx->size - 1) / ctx->size : 0.0;
  b.initialize = [=](size_t n) {
    *count = n;
    bufs->alloc(pickDevice(ctx, o), o.inputs, n * sizeof(T));
    const double stride = static_cast<double>(ctx->size) * o.inputs;
    for (int i = 0; i < o.inputs; i++) {
      cuda::launchFill(bufs->ptrs[i], n, dt, static_cast<double>(ctx->rank * o.inputs + i), stride, bufs->stream);
    }
    GLB_CUDA_CHECK(cudaStreamSynchronize(bufs->stream));
    const auto algo = o.cudaAlgo == "literal" ? named : parseCudaAlgo(o.cudaAlgo, cuda::AllreduceAlgo::AUTO);
    std::vector<cudaStream_t> streams(o.inputs, bufs->stream);
    core->reset(new cuda::CudaAllreduceCore(ctx, bufs->ptrs, n, dt, ReduceOp::SUM, streams, algo, cuda::Workspace::PEER));
  };
  b.run = [=] {
    cudaEventRecord(bufs->e0, bufs->stream);
    (*core)->run();
    cudaEventRecord(bufs->e1, bufs->stream);
    GLB_CUDA_CHECK(cudaStreamSynchronize(bufs->stream));
  };
  b.deviceNs = [=] {
    float ms = 0;
    cudaEventElapsedTime(&ms, bufs->e0, bufs->e1);
    return static_cast<double>(ms) * 1e6;
  };
  b.verify = [=] {
    const size_t n = std::min<size_t>(*count, 1 << 16);
    std::vector<T> host(std::max<size_t>(n, 1));
    std::vector<T*> hp{host.data()};
    for (auto* p : bufs->ptrs) {
      GLB_CUDA_CHECK(cudaMemcpy(host.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
      checkAllreduce<T>(hp, n, ctx->size, o.inputs);
    }
  };
  return b;
}

// Generic CUDA data-movement benchmark over the staged-free registered path.
using CudaBody = std::function<void(cuda::PeerContext&, void* in, const cuda::PeerBuffer& out, size_t n, cudaStream_t)>;

template <typename T>
Benchmark cudaMove(std::shared_ptr<Context> ctx, const Options& o, double busFactor, size_t inMul, size_t outMul,
                   CudaBody body) {
  auto bufs = std::make_shared<CudaBuffers>();
  auto pc = std::make_shared<std::shared_ptr<cuda::PeerContext>>();
  auto out = std::make_shared<std::shared_ptr<cuda::PeerBuffer>>();
  auto count = std::make_shared<size_t>(0);
  Benchmark b;
  b.elementSize = sizeof(T);
  b.busFactor = busFactor;
  b.initialize = [=](size_t n) {
    *count = n;
    const int dev = pickDevice(ctx, o);
    bufs->alloc(dev, 1, std::max<size_t>(1, n * inMul) * sizeof(T));
    *pc = cuda::peerContextFor(ctx, dev);
    *out = (*pc)->allocSymmetric(std::max<size_t>(1, n * outMul) * sizeof(T));
    cuda::launchFill(bufs->ptrs[0], n * inMul, DataTypeOf<T>::value, ctx->rank, 0.0, bufs->stream);
    GLB_CUDA_CHECK(cudaStreamSynchronize(bufs->stream));
  };
  b.run = [=] {
    cudaEventRecord(bufs->e0, bufs->stream);
    body(**pc, bufs->ptrs[0], **out, *count, bufs->stream);
    cudaEventRecord(bufs->e1, bufs->stream);
    GLB_CUDA_CHECK(cudaStreamSynchronize(bufs->stream));
  };
  b.deviceNs = [=] {
    float ms = 0;
    cudaEventElapsedTime(&ms, bufs->e0, bufs->e1);
    return static_cast<double>(ms) * 1e6;
  };
  return b;
}

template <typename T>
void registerTyped(std::map<std::string, BenchmarkFactory>& r, bool half) {
  auto add = [&](const std::string& name, BenchmarkFactory f) {
    // float registers plain names; float16 is selected with --halfprecision at run time.
    r[(half ? "__half__" : "") + name] = std::move(f);
  };
  add("allreduce_ring", [](auto c, const Options& o) { return hostAllreduce<T, AllreduceRing<T>>(c, o); });
  add("allreduce_ring_chunked", [](auto c, const Options& o) { return hostAllreduce<T, AllreduceRingChunked<T>>(c, o); });
  add("allreduce_halving_doubling",
      [](auto c, const Options& o) { return hostAllreduce<T, AllreduceHalvingDoubling<T>>(c, o); });
  add("allreduce_bcube", [](auto c, const Options& o) { return hostAllreduce<T, AllreduceBcube<T>>(c, o); });
  add("allreduce_local", [](auto c, const Options& o) {
    // Reduces this rank's own buffers only (no communication), so the expected value is
    // the sum over the local inputs, whatever the context size.
    auto data = std::make_shared<HostData<T>>();
    auto algo = std::make_shared<std::unique_ptr<Algorithm>>();
    auto count = std::make_shared<size_t>(0);
    const int inputs = o.inputs;
    Benchmark b;
    b.elementSize = sizeof(T);
    b.busFactor = 0.0;
    b.initialize = [=](size_t n) {
      *count = n;
      data->fill(c->rank, c->size, inputs, n);
      algo->reset(new AllreduceLocal<T>(c, data->ptrs, n));
    };
    b.run = [=] { (*algo)->run(); };
    b.verify = [=] {
      const double stride = static_cast<double>(c->size) * inputs;
      for (auto* p : data->ptrs) {
        for (size_t j = 0; j < std::min<size_t>(*count, 1 << 16); j++) {
          const double exp = inputs * (j * stride + c->rank * inputs) + inputs * (inputs - 1) / 2.0;
          const double got = static_cast<double>(static_cast<float>(p[j]));
          const double tol = (sizeof(T) == 2 ? 1e-2 : 1e-5) * std::max(1.0, exp);
          GLB_ENFORCE(std::abs(got - exp) <= tol, "Mismatch at index ", j, ": got ", got, " expected ", exp);
        }
      }
    };
    return b;
  });
  add("new_allreduce", [](auto c, const Options& o) { return newAllreduce<T>(c, o, AllreduceOptions::UNSPECIFIED); });
  add("new_allreduce_ring", [](auto c, const Options& o) { return newAllreduce<T>(c, o, AllreduceOptions::RING); });
  add("new_allreduce_bcube", [](auto c, const Options& o) { return newAllreduce<T>(c, o, AllreduceOptions::BCUBE); });

  add("allgather", [](auto c, const Options&) {
    return simple<T>(c, c->size > 1 ? static_cast<double>(c->size - 1) : 0.0, 1, c->size,
                     [](auto ctx, T* in, T* out, size_t n) {
                       AllgatherOptions o(ctx);
                       o.setInput(in, n);
                       o.setOutput(out, n * ctx->size);
                       allgather(o);
                     });
  });
  add("allgather_v", [](auto c, const Options&) {
    return simple<T>(c, 0.0, 1, c->size, [](auto ctx, T* in, T* out, size_t n) {
      AllgathervOptions o(ctx);
      o.setInput(in, n);
      o.setOutput(out, std::vector<size_t>(ctx->size, n));
      allgatherv(o);
    });
  });
  add("allgather_ring", [](auto c, const Options& o) {
    auto ins = std::make_shared<HostData<T>>();
    auto out = std::make_shared<std::vector<T>>();
    auto algo = std::make_shared<std::unique_ptr<Algorithm>>();
    Benchmark b;
    b.elementSize = sizeof(T);
    b.busFactor = c->size > 1 ? static_cast<double>(c->size - 1) : 0.0;
    const int inputs = o.inputs;
    b.initialize = [=](size_t n) {
      ins->fill(c->rank, c->size, inputs, n);
      out->assign(std::max<size_t>(1, n * inputs * c->size), T(0.0f));
      std::vector<const T*> cp(ins->ptrs.begin(), ins->ptrs.end());
      algo->reset(new AllgatherRing<T>(c, cp, out->data(), n));
    };
    b.run = [=] { (*algo)->run(); };
    return b;
  });
  add("broadcast", [](auto c, const Options&) {
    return simple<T>(c, 1.0, 1, 1, [](auto ctx, T*, T* out, size_t n) {
      BroadcastOptions o(ctx);
      o.setOutput(out, n);
      o.setRoot(0);
      broadcast(o);
    });
  });
  add("broadcast_one_to_all", [](auto c, const Options& o) {
    auto data = std::make_shared<HostData<T>>();
    auto algo = std::make_shared<std::unique_ptr<Algorithm>>();
    Benchmark b;
    b.elementSize = sizeof(T);
    b.busFactor = 1.0;
    const int inputs = o.inputs;
    b.initialize = [=](size_t n) {
      data->fill(c->rank, c->size, inputs, n);
      algo->reset(new BroadcastOneToAll<T>(c, data->ptrs, n));
    };
    b.run = [=] { (*algo)->run(); };
    return b;
  });
  add("reduce", [](auto c, const Options&) {
    return simple<T>(c, 0.0, 1, 1, [](auto ctx, T* in, T* out, size_t n) {
      ReduceOptions o(ctx);
      o.setInput(in, n);
      o.setOutput(out, n);
      o.setRoot(0);
      o.setReduceFunction([](void* cc, const void* a, const void* bb, size_t k) { sum<T>(cc, a, bb, k); });
      reduce(o);
    });
  });
  add("reduce_scatter", [](auto c, const Options& o) {
    auto data = std::make_shared<HostData<T>>();
    auto algo = std::make_shared<std::unique_ptr<Algorithm>>();
    Benchmark b;
    b.elementSize = sizeof(T);
    b.busFactor = c->size > 1 ? static_cast<double>(c->size - 1) / c->size : 0.0;
    const int inputs = o.inputs;
    b.initialize = [=](size_t n) {
      data->fill(c->rank, c->size, inputs, n);
      std::vector<int> recv;
      size_t left = n, per = (n + c->size - 1) / c->size;
      for (int i = 0; i < c->size; i++) {
        size_t k = std::min(per, left);
        recv.push_back(static_cast<int>(k));
        left -= k;
      }
      algo->reset(new ReduceScatterHalvingDoubling<T>(c, data->ptrs, n, recv));
    };
    b.run = [=] { (*algo)->run(); };
    return b;
  });
  add("new_reduce_scatter", [](auto c, const Options&) {
    return simple<T>(c, c->size > 1 ? static_cast<double>(c->size - 1) / c->size : 0.0, 1, 1,
                     [](auto ctx, T* in, T* out, size_t n) {
                       ReduceScatterOptions o(ctx);
                       o.setInput(in, n);
                       o.setOutput(out, detail::subRange({0, n}, ctx->size, ctx->rank).len);
                       o.setReduceFunction([](void* cc, const void* a, const void* bb, size_t k) { sum<T>(cc, a, bb, k); });
                       reduce_scatter(o);
                     });
  });
  add("scatter", [](auto c, const Options&) {
    return simple<T>(c, 0.0, c->size, 1, [](auto ctx, T* in, T* out, size_t n) {
      ScatterOptions o(ctx);
      if (ctx->rank == 0) {
        std::vector<T*> ins;
        for (int i = 0; i < ctx->size; i++) ins.push_back(in + i * n);
        o.setInputs(ins, n);
      }
      o.setOutput(out, n);
      o.setRoot(0);
      scatter(o);
    });
  });
  add("gather", [](auto c, const Options&) {
    return simple<T>(c, 0.0, 1, c->size, [](auto ctx, T* in, T* out, size_t n) {
      GatherOptions o(ctx);
      o.setInput(in, n);
      if (ctx->rank == 0) o.setOutput(out, n * ctx->size);
      o.setRoot(0);
      gather(o);
    });
  });

  // ---- CUDA --------------------------------------------------------------------------
  add("cuda_allreduce_ring", [](auto c, const Options& o) { return cudaAllreduce<T>(c, o, cuda::AllreduceAlgo::RING); });
  add("cuda_allreduce_ring_chunked",
      [](auto c, const Options& o) { return cudaAllreduce<T>(c, o, cuda::AllreduceAlgo::RING_CHUNKED); });
  add("cuda_allreduce_halving_doubling",
      [](auto c, const Options& o) { return cudaAllreduce<T>(c, o, cuda::AllreduceAlgo::HALVING_DOUBLING); });
  add("cuda_allreduce_halving_doubling_pipelined",
      [](auto c, const Options& o) { return cudaAllreduce<T>(c, o, cuda::AllreduceAlgo::HALVING_DOUBLING_PIPELINED); });
  add("cuda_allreduce_bcube", [](auto c, const Options& o) { return cudaAllreduce<T>(c, o, cuda::AllreduceAlgo::BCUBE); });
  add("cuda_broadcast_one_to_all", [](auto c, const Options& o) {
    return cudaMove<T>(c, o, 1.0, 1, 1, [](cuda::PeerContext& pc, void*, const cuda::PeerBuffer& out, size_t n, cudaStream_t s) {
      cuda::broadcast(pc, out, 0, n * sizeof(T), 0, s);
    });
  });
  add("cuda_allgather", [](auto c, const Options& o) {
    return cudaMove<T>(c, o, c->size > 1 ? static_cast<double>(c->size - 1) : 0.0, 1, c->size,
                       [](cuda::PeerContext& pc, void* in, const cuda::PeerBuffer& out, size_t n, cudaStream_t s) {
                         cuda::allgatherv(pc, in, out, 0, std::vector<size_t>(pc.size, n * sizeof(T)), s);
                       });
  });
  add("cuda_alltoall", [](auto c, const Options& o) {
    return cudaMove<T>(c, o, c->size > 1 ? static_cast<double>(c->size - 1) : 0.0, c->size, c->size,
                       [](cuda::PeerContext& pc, void* in, const cuda::PeerBuffer& out, size_t n, cudaStream_t s) {
                         std::vector<size_t> b(pc.size, n * sizeof(T));
                         cuda::alltoallv(pc, in, b, out, 0, b, s);
                       });
  });
  // Ring rotation (every rank sends n elements to rank+1 and receives from rank-1): through the
  // mailbox ring into a plain pointer, and zero-copy into the symmetric buffer.
  add("cuda_sendrecv", [](auto c, const Options& o) {
    return cudaMove<T>(c, o, c->size > 1 ? 1.0 : 0.0, 1, 1,
                       [](cuda::PeerContext& pc, void* in, const cuda::PeerBuffer& out, size_t n, cudaStream_t s) {
                         if (pc.size == 1) return;
                         cuda::sendrecv(pc, in, n * sizeof(T), (pc.rank + 1) % pc.size, out.local, n * sizeof(T),
                                        (pc.rank + pc.size - 1) % pc.size, s);
                       });
  });
  add("cuda_exchange", [](auto c, const Options& o) {
    return cudaMove<T>(c, o, c->size > 1 ? 1.0 : 0.0, 1, 1,
                       [](cuda::PeerContext& pc, void* in, const cuda::PeerBuffer& out, size_t n, cudaStream_t s) {
                         if (pc.size == 1) return;
                         cuda::exchange(pc, in, n * sizeof(T), (pc.rank + 1) % pc.size, out, 0, n * sizeof(T),
                                        (pc.rank + pc.size - 1) % pc.size, s);
                       });
  });
  add("cuda_reduce_scatter", [](auto c, const Options& o) {
    // `n` is the full input length (as in the reference's reduce_scatter benchmark).
    auto inBuf = std::make_shared<std::shared_ptr<cuda::PeerBuffer>>();
    return cudaMove<T>(c, o, c->size > 1 ? static_cast<double>(c->size - 1) / c->size : 0.0, 1, 1,
                       [inBuf](cuda::PeerContext& pc, void* in, const cuda::PeerBuffer& out, size_t n, cudaStream_t s) {
                         if (!*inBuf) *inBuf = pc.registerBuffer(in, std::max<size_t>(1, n) * sizeof(T));
                         std::vector<size_t> counts(pc.size);
                         for (int r = 0; r < pc.size; r++) counts[r] = detail::subRange({0, n}, pc.size, r).len;
                         cuda::reduce_scatter(pc, **inBuf, 0, out.local, counts, DataTypeOf<T>::value, ReduceOp::SUM, s);
                       });
  });
}

}  // namespace

const std::map<std::string, BenchmarkFactory>& benchmarkRegistry() {
  static std::map<std::string, BenchmarkFactory> reg = [] {
    std::map<std::string, BenchmarkFactory> r;
    std::map<std::string, BenchmarkFactory> f32, f16;
    registerTyped<float>(f32, false);
    registerTyped<float16>(f16, true);
    for (auto& kv : f32) {
      const std::string name = kv.first;
      BenchmarkFactory fl = kv.second;
      BenchmarkFactory hf = f16["__half__" + name];
      r[name] = [fl, hf](std::shared_ptr<Context> c, const Options& o) { return o.halfPrecision ? hf(c, o) : fl(c, o); };
    }
    // ---- untyped benchmarks -----------------------------------------------------------
    r["barrier_all_to_all"] = [](std::shared_ptr<Context> c, const Options&) {
      auto a = std::make_shared<BarrierAllToAll>(c);
      Benchmark b;
      b.initialize = [](size_t) {};
      b.run = [a] { a->run(); };
      return b;
    };
    r["barrier_all_to_one"] = [](std::shared_ptr<Context> c, const Options&) {
      auto a = std::make_shared<BarrierAllToOne>(c);
      Benchmark b;
      b.initialize = [](size_t) {};
      b.run = [a] { a->run(); };
      return b;
    };
    r["barrier"] = [](std::shared_ptr<Context> c, const Options&) {
      Benchmark b;
      b.initialize = [](size_t) {};
      b.run = [c] {
        BarrierOptions o(c);
        barrier(o);
      };
      return b;
    };
    r["pairwise_exchange"] = [](std::shared_ptr<Context> c, const Options& o) {
      auto a = std::make_shared<std::unique_ptr<PairwiseExchange>>();
      Benchmark b;
      b.elementSize = 1;
      const int dest = o.destinations;
      b.initialize = [=](size_t n) { a->reset(new PairwiseExchange(c, static_cast<int>(n), dest)); };
      b.run = [a] { (*a)->run(); };
      return b;
    };
    r["alltoall"] = [](std::shared_ptr<Context> c, const Options&) {
      return simple<uint64_t>(c, c->size > 1 ? static_cast<double>(c->size - 1) : 0.0, c->size, c->size,
                              [](auto ctx, uint64_t* in, uint64_t* out, size_t n) {
                                AlltoallOptions o(ctx);
                                o.setInput(in, n * ctx->size);
                                o.setOutput(out, n * ctx->size);
                                alltoall(o);
                              });
    };
    r["alltoall_v"] = [](std::shared_ptr<Context> c, const Options&) {
      // rank r sends n*(r + P - i) elements to rank i (main.cc:412-415).
      const int P = c->size;
      return simple<uint64_t>(c, 0.0, static_cast<size_t>(2 * P) * P, static_cast<size_t>(2 * P) * P,
                              [](auto ctx, uint64_t* in, uint64_t* out, size_t n) {
                                const int P = ctx->size, r = ctx->rank;
                                std::vector<int64_t> send(P), recv(P);
                                for (int i = 0; i < P; i++) {
                                  send[i] = static_cast<int64_t>(n) * (r + P - i);
                                  recv[i] = static_cast<int64_t>(n) * (i + P - r);
                                }
                                AlltoallvOptions o(ctx);
                                o.setInput(in, send);
                                o.setOutput(out, recv);
                                alltoallv(o);
                              });
    };
    auto sendrecv = [](bool roundtrip, bool nonblocking) {
      return [roundtrip, nonblocking](std::shared_ptr<Context> c, const Options& o) {
        GLB_ENFORCE_EQ(c->size, 2, "send/recv benchmarks need exactly two processes");
        auto buf = std::make_shared<std::vector<float>>();
        auto ub = std::make_shared<std::unique_ptr<transport::UnboundBuffer>>();
        Benchmark b;
        const int messages = o.messages;
        b.initialize = [=](size_t n) {
          buf->assign(std::max<size_t>(1, n), 1.0f);
          *ub = c->createUnboundBuffer(buf->data(), n * sizeof(float));
        };
        b.run = [=] {
          constexpr uint64_t kSlot = 0x1337;
          auto& u = *ub;
          if (roundtrip) {
            if (c->rank == 0) {
              u->send(1, kSlot);
              u->waitSend();
              u->recv(1, kSlot);
              u->waitRecv();
            } else {
              u->recv(0, kSlot);
              u->waitRecv();
              u->send(0, kSlot);
              u->waitSend();
            }
            return;
          }
          if (nonblocking) {
            for (int i = 0; i < messages; i++) (c->rank == 0) ? u->send(1, kSlot) : u->recv(0, kSlot);
            for (int i = 0; i < messages; i++) (c->rank == 0) ? u->waitSend() : u->waitRecv();
          } else {
            for (int i = 0; i < messages; i++) {
              if (c->rank == 0) {
                u->send(1, kSlot);
                u->waitSend();
              } else {
                u->recv(0, kSlot);
                u->waitRecv();
              }
            }
          }
        };
        return b;
      };
    };
    r["sendrecv_roundtrip"] = sendrecv(true, false);
    r["sendrecv_stress"] = sendrecv(false, false);
    r["isendirecv_stress"] = sendrecv(false, true);
    r["cuda_barrier"] = [](std::shared_ptr<Context> c, const Options& o) {
      auto pc = std::make_shared<std::shared_ptr<cuda::PeerContext>>();
      Benchmark b;
      b.initialize = [=](size_t) { *pc = cuda::peerContextFor(c, pickDevice(c, o)); };
      b.run = [=] {
        cuda::barrier(**pc, nullptr);
        cudaStreamSynchronize(nullptr);
      };
      return b;
    };
    return r;
  }();
  return reg;
}

}  // namespace benchmark
}  // namespace glb
