// Fused allreduce kernels over NVLink peer memory (sm_100a).
//
//   barrierKernel        flag barrier only.
//   oneShotAllreduce     latency regime. Each rank stages its input in its own
//                        symmetric pool (double-buffered by launch parity), one
//                        flag barrier, then every rank reads all P staged copies
//                        straight over NVLink, reduces in registers (fp32
//                        accumulate for 16-bit types) and writes its output. One
//                        kernel, one barrier, no host involvement; works for any
//                        user pointer because peers only touch the pools.
//   twoShotAllreduce     bandwidth regime, in place on peer-mapped user buffers.
//                        Rank r owns chunk r: it loads that chunk from all P
//                        buffers (reduce-scatter by direct peer reads), reduces,
//                        and immediately stores the result into all P buffers
//                        (allgather by direct peer writes) — reduce, scatter and
//                        gather fused in one pass; 2·S·(P-1)/P bytes per GPU in
//                        two hops instead of the ring's 2(P-1).
//   nvlsAllreduce        same ownership, but the reduction happens inside the
//                        NVSwitch: multimem.ld_reduce pulls the switch-reduced
//                        chunk, multimem.st broadcasts it; ~S(1+1/P) per direction.
//
// There is no reference counterpart: the reference stages through pinned host
// memory + TCP (cuda_allreduce_ring_chunked.cc:129-273) and reduces on the CPU.
#include "glb/cuda/device_common.cuh"
#include "glb/cuda/kernels.h"

namespace glb {
namespace cuda {

bool oneShotPushEnabled();

__global__ void barrierKernel(CommArgs a) {
  const uint32_t e = loadEpoch(a);
  blockBarrier(a, e + 1);
  retire(a, 1, 0);
}

// ---- one-shot ---------------------------------------------------------------------

// Push flavour for the smallest messages: every rank STORES its contribution into slot
// `rank` of every peer's pool (posted writes, no round trip), one barrier, then reads the P
// slots from its own memory. Saves the remote-load round trip of the pull flavour at the
// cost of P x the staging space, so it is used while P * bytes fits one half.
template <typename T>
__global__ void __launch_bounds__(kThreads)
oneShotPushAllreduceKernel(CommArgs a, const T* in, T* out, size_t count, DevOp op, PeerPtrs stage,
                           size_t halfBytes, size_t slotBytes, bool vectorOk) {
  using PT = PackTraits<T>;
  const uint32_t e = loadEpoch(a);
  const uint32_t parity = ld_relaxed_sys(&a.sig[a.rank]->stageSeq) & 1u;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t base = parity * halfBytes;
  const size_t nvec = vectorOk ? count / PT::kElems : 0;
  const size_t tailStart = nvec * PT::kElems;
  const size_t mySlot = base + static_cast<size_t>(a.rank) * slotBytes;

  for (size_t v = tid; v < nvec; v += nthreads) {
    const Pack16 p = ld128_stream(reinterpret_cast<const char*>(in) + v * 16);
#pragma unroll
    for (int r = 0; r < kMaxRanks; r++) {
      if (r < a.nranks) st128_stream(static_cast<char*>(stage.p[r]) + mySlot + v * 16, p);
    }
  }
  for (size_t i = tailStart + tid; i < count; i += nthreads) {
    const T x = in[i];
    for (int r = 0; r < a.nranks; r++) reinterpret_cast<T*>(static_cast<char*>(stage.p[r]) + mySlot)[i] = x;
  }

  blockBarrier(a, e + 1);

  const char* mine = static_cast<const char*>(stage.p[a.rank]) + base;
  for (size_t v = tid; v < nvec; v += nthreads) {
    Pack16 p[kMaxRanks];
#pragma unroll
    for (int r = 0; r < kMaxRanks; r++) {
      if (r < a.nranks) p[r] = ld128(mine + r * slotBytes + v * 16);
    }
    typename PT::AccPack acc = PT::widen(p[0]);
#pragma unroll
    for (int r = 1; r < kMaxRanks; r++) {
      if (r < a.nranks) PT::combine(acc, p[r], op);
    }
    st128(reinterpret_cast<char*>(out) + v * 16, PT::narrow(acc));
  }
  for (size_t i = tailStart + tid; i < count; i += nthreads) {
    T acc = reinterpret_cast<const T*>(mine)[i];
    for (int r = 1; r < a.nranks; r++) acc = PT::combineOne(acc, reinterpret_cast<const T*>(mine + r * slotBytes)[i], op);
    out[i] = acc;
  }
  retire(a, 1, 1);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
oneShotAllreduceKernel(CommArgs a, const T* in, T* out, size_t count, DevOp op,
                       PeerPtrs stage, size_t halfBytes, bool vectorOk) {
  using PT = PackTraits<T>;
  const uint32_t e = loadEpoch(a);
  const uint32_t parity = ld_relaxed_sys(&a.sig[a.rank]->stageSeq) & 1u;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t stageOff = parity * halfBytes;
  char* myStage = static_cast<char*>(stage.p[a.rank]) + stageOff;

  const size_t nvec = vectorOk ? count / PT::kElems : 0;
  const size_t tailStart = nvec * PT::kElems;

  // Phase 0: publish my contribution in my pool.
  for (size_t v = tid; v < nvec; v += nthreads) {
    st128(myStage + v * 16, ld128_stream(reinterpret_cast<const char*>(in) + v * 16));
  }
  for (size_t i = tailStart + tid; i < count; i += nthreads) {
    reinterpret_cast<T*>(myStage)[i] = in[i];
  }

  blockBarrier(a, e + 1);

  // Phase 1: reduce all P staged copies. Summation order is rank order on every
  // rank, so all ranks produce bit-identical results.
  for (size_t v = tid; v < nvec; v += nthreads) {
    Pack16 p[kMaxRanks];
#pragma unroll
    for (int r = 0; r < kMaxRanks; r++) {
      if (r < a.nranks) p[r] = ld128_stream(static_cast<const char*>(stage.p[r]) + stageOff + v * 16);
    }
    typename PT::AccPack acc = PT::widen(p[0]);
#pragma unroll
    for (int r = 1; r < kMaxRanks; r++) {
      if (r < a.nranks) PT::combine(acc, p[r], op);
    }
    st128(reinterpret_cast<char*>(out) + v * 16, PT::narrow(acc));
  }
  for (size_t i = tailStart + tid; i < count; i += nthreads) {
    T acc = reinterpret_cast<const T*>(static_cast<const char*>(stage.p[0]) + stageOff)[i];
    for (int r = 1; r < a.nranks; r++) {
      acc = PT::combineOne(acc, reinterpret_cast<const T*>(static_cast<const char*>(stage.p[r]) + stageOff)[i], op);
    }
    out[i] = acc;
  }
  retire(a, 1, 1);
}

// ---- two-shot (fused reduce-scatter + allgather, in place) --------------------------

// [begin, end) of rank r's share of n items.
__device__ __forceinline__ void shareOf(size_t n, int parts, int r, size_t& begin, size_t& end) {
  const size_t base = n / parts, rem = n % parts;
  begin = r * base + (static_cast<size_t>(r) < rem ? r : rem);
  end = begin + base + (static_cast<size_t>(r) < rem ? 1 : 0);
}

template <typename T, int NR, int UNROLL>
__global__ void __launch_bounds__(kThreads)
twoShotAllreduceKernel(CommArgs a, PeerPtrs bufs, size_t count, DevOp op, bool vectorOk) {
  using PT = PackTraits<T>;
  const int P = NR > 0 ? NR : a.nranks;
  const uint32_t e = loadEpoch(a);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;

  // Everybody's kernel has started => everybody's input is final.
  blockBarrier<false>(a, e + 1);

  const size_t nvec = vectorOk ? count / PT::kElems : 0;
  size_t vb, ve;
  shareOf(nvec, P, a.rank, vb, ve);

  // Peer order is rotated by rank so that at any instant the P readers hit P
  // different sources. Rank r alone computes chunk r, so the (rotated) summation
  // order cannot make ranks disagree.
  constexpr int kSlots = NR > 0 ? NR : kMaxRanks;
  char* peer[kSlots];
#pragma unroll
  for (int i = 0; i < kSlots; i++) peer[i] = i < P ? static_cast<char*>(bufs.p[(a.rank + i) % P]) : nullptr;

  for (size_t v0 = vb + tid; v0 < ve; v0 += nthreads * UNROLL) {
    Pack16 p[UNROLL][kSlots];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t v = v0 + static_cast<size_t>(u) * nthreads;
      if (v < ve) {
#pragma unroll
        for (int i = 0; i < kSlots; i++) {
          if (i < P) p[u][i] = ld128_stream(peer[i] + v * 16);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t v = v0 + static_cast<size_t>(u) * nthreads;
      if (v < ve) {
        typename PT::AccPack acc = PT::widen(p[u][0]);
#pragma unroll
        for (int i = 1; i < kSlots; i++) {
          if (i < P) PT::combine(acc, p[u][i], op);
        }
        const Pack16 res = PT::narrow(acc);
#pragma unroll
        for (int i = 0; i < kSlots; i++) {
          if (i < P) st128_stream(peer[i] + v * 16, res);
        }
      }
    }
  }

  // Scalar tail (count not a multiple of the pack width, or unaligned buffers):
  // split across ranks the same way.
  {
    const size_t tailStart = nvec * PT::kElems;
    size_t tb, te;
    shareOf(count - tailStart, P, a.rank, tb, te);
    for (size_t i = tailStart + tb + tid; i < tailStart + te; i += nthreads) {
      T acc = static_cast<const T*>(bufs.p[0])[i];
      for (int r = 1; r < P; r++) acc = PT::combineOne(acc, static_cast<const T*>(bufs.p[r])[i], op);
      for (int r = 0; r < P; r++) static_cast<T*>(bufs.p[r])[i] = acc;
    }
  }

  // All my stores have landed everywhere and nobody still reads my buffer.
  blockBarrier(a, e + 2);
  retire(a, 2, 0);
}

// ---- NVLS -----------------------------------------------------------------------------

template <typename T, int UNROLL>
__global__ void __launch_bounds__(kThreads)
nvlsAllreduceKernel(CommArgs a, char* mcBase, PeerPtrs bufs, size_t count) {
  using PT = PackTraits<T>;
  const uint32_t e = loadEpoch(a);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  blockBarrier<false>(a, e + 1);
  const size_t nvec = count / PT::kElems;
  size_t vb, ve;
  shareOf(nvec, a.nranks, a.rank, vb, ve);
  for (size_t v0 = vb + tid; v0 < ve; v0 += nthreads * UNROLL) {
    Pack16 r[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t v = v0 + static_cast<size_t>(u) * nthreads;
      if (v < ve) r[u] = Multimem<T>::ldReduceAdd(mcBase + v * 16);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const size_t v = v0 + static_cast<size_t>(u) * nthreads;
      if (v < ve) multimemSt128(mcBase + v * 16, r[u]);
    }
  }
  // Sub-pack tail through plain peer pointers (last rank).
  if (a.rank == a.nranks - 1) {
    for (size_t i = nvec * PT::kElems + tid; i < count; i += nthreads) {
      T acc = static_cast<const T*>(bufs.p[0])[i];
      for (int r = 1; r < a.nranks; r++) acc = PT::combineOne(acc, static_cast<const T*>(bufs.p[r])[i], DevOp::SUM);
      for (int r = 0; r < a.nranks; r++) static_cast<T*>(bufs.p[r])[i] = acc;
    }
  }
  blockBarrier(a, e + 2);
  retire(a, 2, 0);
}

// ---- host-side launchers ----------------------------------------------------------------

namespace {

template <typename F>
void dispatchType(DataType dt, F&& f) {
  switch (dt) {
    case DataType::INT8: f(int8_t{}); break;
    case DataType::UINT8: f(uint8_t{}); break;
    case DataType::INT16: f(int16_t{}); break;
    case DataType::INT32: f(int32_t{}); break;
    case DataType::UINT32: f(uint32_t{}); break;
    case DataType::INT64: f((long long){}); break;
    case DataType::UINT64: f((unsigned long long){}); break;
    case DataType::FLOAT32: f(float{}); break;
    case DataType::FLOAT64: f(double{}); break;
    case DataType::FLOAT16: f(__half{}); break;
    case DataType::BFLOAT16: f(__nv_bfloat16{}); break;
  }
}

template <typename T>
constexpr bool isHotType() {
  return std::is_same<T, float>::value || std::is_same<T, __half>::value || std::is_same<T, __nv_bfloat16>::value;
}

}  // namespace

// Force-load every kernel of this file. With CUDA's lazy module loading the first
// launch of a kernel may have to synchronise the context; if a peer rank on the same
// device is already spinning inside its collective kernel that launch never happens
// (documented lazy-loading deadlock for kernels that assume concurrency).
namespace {
template <typename K>
void touch(K kernel) {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(kernel));
}
}  // namespace

void preloadAllreduceKernels() {
  touch(barrierKernel);
  for (DataType dt : {DataType::INT8, DataType::UINT8, DataType::INT16, DataType::INT32, DataType::UINT32, DataType::INT64,
                      DataType::UINT64, DataType::FLOAT32, DataType::FLOAT64, DataType::FLOAT16, DataType::BFLOAT16}) {
    dispatchType(dt, [&](auto tag) {
      using T = decltype(tag);
      touch(oneShotAllreduceKernel<T>);
      touch(oneShotPushAllreduceKernel<T>);
      touch(twoShotAllreduceKernel<T, 0, 1>);
      if constexpr (isHotType<T>()) {
        touch(twoShotAllreduceKernel<T, 2, 4>);
        touch(twoShotAllreduceKernel<T, 4, 2>);
        touch(twoShotAllreduceKernel<T, 8, 2>);
        touch(nvlsAllreduceKernel<T, 4>);
      }
    });
  }
  cudaGetLastError();
}

bool& oneShotPushFlag() {
  static bool v = true;
  return v;
}
bool oneShotPushEnabled() { return oneShotPushFlag(); }
void setOneShotPush(bool on) { oneShotPushFlag() = on; }

void launchBarrier(const CommArgs& a, cudaStream_t stream) {
  barrierKernel<<<1, 32, 0, stream>>>(a);
}

void launchOneShotAllreduce(const CommArgs& a, const void* in, void* out, size_t count, DataType dt, ReduceOp op,
                            const PeerPtrs& stage, size_t halfBytes, int blocks, cudaStream_t stream) {
  const bool vectorOk = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  // Push while every rank's copy fits a slot (half / P), else pull.
  const size_t slotBytes = (halfBytes / static_cast<size_t>(a.nranks)) / 16 * 16;
  const bool push = count * elementSize(dt) <= slotBytes && oneShotPushEnabled();
  dispatchType(dt, [&](auto tag) {
    using T = decltype(tag);
    if (push) {
      oneShotPushAllreduceKernel<T><<<blocks, kThreads, 0, stream>>>(a, static_cast<const T*>(in), static_cast<T*>(out),
                                                                     count, static_cast<DevOp>(op), stage, halfBytes,
                                                                     slotBytes, vectorOk);
    } else {
      oneShotAllreduceKernel<T><<<blocks, kThreads, 0, stream>>>(a, static_cast<const T*>(in), static_cast<T*>(out),
                                                                 count, static_cast<DevOp>(op), stage, halfBytes,
                                                                 vectorOk);
    }
  });
}

void launchTwoShotAllreduce(const CommArgs& a, const PeerPtrs& bufs, size_t count, DataType dt, ReduceOp op,
                            bool vectorOk, int blocks, cudaStream_t stream) {
  const DevOp dop = static_cast<DevOp>(op);
  dispatchType(dt, [&](auto tag) {
    using T = decltype(tag);
    if constexpr (isHotType<T>()) {
      switch (a.nranks) {
        case 2: twoShotAllreduceKernel<T, 2, 4><<<blocks, kThreads, 0, stream>>>(a, bufs, count, dop, vectorOk); return;
        case 4: twoShotAllreduceKernel<T, 4, 2><<<blocks, kThreads, 0, stream>>>(a, bufs, count, dop, vectorOk); return;
        case 8: twoShotAllreduceKernel<T, 8, 2><<<blocks, kThreads, 0, stream>>>(a, bufs, count, dop, vectorOk); return;
        default: break;
      }
    }
    twoShotAllreduceKernel<T, 0, 1><<<blocks, kThreads, 0, stream>>>(a, bufs, count, dop, vectorOk);
  });
}

bool nvlsSupports(DataType dt, ReduceOp op) {
  return op == ReduceOp::SUM && (dt == DataType::FLOAT32 || dt == DataType::FLOAT16 || dt == DataType::BFLOAT16);
}

void launchNvlsAllreduce(const CommArgs& a, void* mcPtr, const PeerPtrs& bufs, size_t count, DataType dt, int blocks,
                         cudaStream_t stream) {
  char* mc = static_cast<char*>(mcPtr);
  switch (dt) {
    case DataType::FLOAT32: nvlsAllreduceKernel<float, 4><<<blocks, kThreads, 0, stream>>>(a, mc, bufs, count); break;
    case DataType::FLOAT16: nvlsAllreduceKernel<__half, 4><<<blocks, kThreads, 0, stream>>>(a, mc, bufs, count); break;
    case DataType::BFLOAT16: nvlsAllreduceKernel<__nv_bfloat16, 4><<<blocks, kThreads, 0, stream>>>(a, mc, bufs, count); break;
    default: break;
  }
}

}  // namespace cuda
}  // namespace glb
