// Pipelined allreduce for ARBITRARY device pointers (sm_100a).
//
// The reference API takes plain `T*` (cuda_allreduce_ring_chunked.h:30-36); a pointer that
// is neither registered nor symmetric cannot be read by the peers (and cannot be bound to an
// NVSwitch multicast object), so the data has to pass through the symmetric pool. Instead
// of copy-in / collective / copy-out as three passes, ONE persistent kernel runs a
// three-stage software pipeline over chunks of the vector, one device barrier per step:
//
//     step i:   A(i)    copy chunk i        user input  -> pool slot i % 3      (local HBM)
//               B(i-1)  reduce chunk i-1    in the pool: multimem.ld_reduce + multimem.st
//                                           through the NVSwitch (or peer loads + stores)
//               C(i-2)  copy chunk i-2      pool slot -> user output              (local HBM)
//
// The three stages of a step touch three different slots, so they run concurrently: the
// warps of every CTA are split into an exchange group (B, NVLink-bound) and a copy group
// (A and C, HBM-bound) and the local copies hide under the NVLink time.
//
// Ownership: a chunk is P shares x G tiles x T 16-byte groups; CTA b owns tile b of every
// share in every stage, so the CTA-to-CTA flag barrier is all the synchronisation needed
// (the CTA that reads a peer's tile is barrier-paired with the CTA that wrote it).
#include "glb/cuda/device_common.cuh"
#include "glb/cuda/kernels.h"

namespace glb {
namespace cuda {

namespace {

template <typename T>
__device__ __forceinline__ Pack16 loadGroupP(const T* base, size_t g, size_t count, bool aligned) {
  constexpr int K = 16 / sizeof(T);
  const size_t i0 = g * K;
  if (aligned && i0 + K <= count) return ld128_stream(reinterpret_cast<const char*>(base) + g * 16);
  Pack16 p;
  p.w[0] = p.w[1] = p.w[2] = p.w[3] = 0u;
  T* t = reinterpret_cast<T*>(&p);
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (i0 + k < count) t[k] = base[i0 + k];
  }
  return p;
}

template <typename T>
__device__ __forceinline__ void storeGroupP(T* base, size_t g, size_t count, bool aligned, const Pack16& p) {
  constexpr int K = 16 / sizeof(T);
  const size_t i0 = g * K;
  if (aligned && i0 + K <= count) {
    st128_stream(reinterpret_cast<char*>(base) + g * 16, p);
    return;
  }
  const T* t = reinterpret_cast<const T*>(&p);
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (i0 + k < count) base[i0 + k] = t[k];
  }
}

}  // namespace

template <typename T, bool MC, int NR>
__global__ void __launch_bounds__(kThreads)
pipelinedAllreduceKernel(CommArgs a, const T* in, T* out, size_t count, DevOp op, float scale, PeerPtrs stage,
                         char* mcStage, int tileVecs, int exchangeThreads, LocalPtrs extra) {
  using PT = PackTraits<T>;
  const int P = NR > 0 ? NR : a.nranks;
  const int G = gridDim.x;
  const int b = blockIdx.x;
  const size_t T_ = static_cast<size_t>(tileVecs);  // power of two (host enforces)
  const int tileShift = 31 - __clz(tileVecs);
  const size_t chunkVecs = static_cast<size_t>(P) * G * T_;
  const size_t slotBytes = chunkVecs * 16;
  const size_t ngroups = (count + PT::kElems - 1) / PT::kElems;
  const size_t nchunks = (ngroups + chunkVecs - 1) / chunkVecs;
  const uint32_t e = loadEpoch(a);
  char* myStage = static_cast<char*>(stage.p[a.rank]);

  bool aligned = reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  for (int k = 0; k < extra.n; k++) aligned = aligned && reinterpret_cast<uintptr_t>(extra.p[k]) % 16 == 0;

  const bool exchanger = static_cast<int>(threadIdx.x) < exchangeThreads;
  const int xt = threadIdx.x, xn = exchangeThreads;                            // exchange group
  const int ct = threadIdx.x - exchangeThreads, cn = blockDim.x - exchangeThreads;  // copy group

  uint32_t used = 0;
  for (size_t i = 0; i < nchunks + 2; i++) {
    if (exchanger) {
      // ---- B(i-1): reduce my share's tile of chunk i-1 in the pool, deliver to everyone ----
      if (i >= 1 && i - 1 < nchunks) {
        const size_t slot = ((i - 1) % 3) * slotBytes;
        const size_t w0 = (static_cast<size_t>(a.rank) * G + b) * T_;
        if constexpr (MC) {
          char* mc = mcStage + slot + w0 * 16;
          constexpr int U = 4;
          for (size_t o0 = xt; o0 < T_; o0 += static_cast<size_t>(xn) * U) {
            Pack16 r[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              const size_t o = o0 + static_cast<size_t>(u) * xn;
              if (o < T_) r[u] = Multimem<T>::ldReduceAdd(mc + o * 16);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
              const size_t o = o0 + static_cast<size_t>(u) * xn;
              if (o < T_) {
                if (scale != 1.0f) {
                  typename PT::AccPack acc = PT::widen(r[u]);
                  PT::scale(acc, scale);
                  r[u] = PT::narrow(acc);
                }
                multimemSt128(mc + o * 16, r[u]);
              }
            }
          }
        } else {
          // Peer exchange inside the pool, like the two-shot kernel: kSlots x U independent
          // 128-bit loads in flight per thread (an NVLink round trip is ~2 us).
          constexpr int kSlots = NR > 0 ? NR : kMaxRanks;
          constexpr int U = NR == 2 ? 8 : (NR == 4 ? 4 : (NR == 8 ? 2 : 1));
          char* peer[kSlots];
#pragma unroll
          for (int q = 0; q < kSlots; q++) {
            peer[q] = q < P ? static_cast<char*>(stage.p[(a.rank + q) % P]) + slot + w0 * 16 : nullptr;
          }
          for (size_t o0 = xt; o0 < T_; o0 += static_cast<size_t>(xn) * U) {
            Pack16 v[U][kSlots];
#pragma unroll
            for (int u = 0; u < U; u++) {
              const size_t o = o0 + static_cast<size_t>(u) * xn;
              if (o < T_) {
#pragma unroll
                for (int q = 0; q < kSlots; q++) {
                  if (q < P) v[u][q] = ld128_stream(peer[q] + o * 16);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
              const size_t o = o0 + static_cast<size_t>(u) * xn;
              if (o < T_) {
                typename PT::AccPack acc = PT::widen(v[u][0]);
#pragma unroll
                for (int q = 1; q < kSlots; q++) {
                  if (q < P) PT::combine(acc, v[u][q], op);
                }
                if (scale != 1.0f) PT::scale(acc, scale);
                const Pack16 res = PT::narrow(acc);
#pragma unroll
                for (int q = 0; q < kSlots; q++) {
                  if (q < P) st128_stream(peer[q] + o * 16, res);
                }
              }
            }
          }
        }
      }
    } else {
      constexpr int CU = 4;  // independent 128-bit accesses in flight per copy thread
      const size_t PT_ = static_cast<size_t>(P) * T_;
      // ---- C(i-2): pool -> user output ----------------------------------------------------
      if (i >= 2) {
        const size_t c = i - 2;
        const char* slot = myStage + (c % 3) * slotBytes;
        const size_t g0 = c * chunkVecs;
        for (size_t j0 = ct; j0 < PT_; j0 += static_cast<size_t>(cn) * CU) {
          Pack16 p[CU];
          size_t g[CU];
#pragma unroll
          for (int u = 0; u < CU; u++) {
            const size_t j = j0 + static_cast<size_t>(u) * cn;
            const size_t w = (((j >> tileShift) * G + b) << tileShift) + (j & (T_ - 1));
            g[u] = j < PT_ ? g0 + w : ngroups;
            if (g[u] < ngroups) p[u] = ld128(slot + w * 16);
          }
#pragma unroll
          for (int u = 0; u < CU; u++) {
            if (g[u] < ngroups) {
              storeGroupP(out, g[u], count, aligned, p[u]);
              for (int k = 0; k < extra.n; k++) storeGroupP(static_cast<T*>(extra.p[k]), g[u], count, aligned, p[u]);
            }
          }
        }
      }
      // ---- A(i): user input -> pool -------------------------------------------------------
      if (i < nchunks) {
        char* slot = myStage + (i % 3) * slotBytes;
        const size_t g0 = i * chunkVecs;
        for (size_t j0 = ct; j0 < PT_; j0 += static_cast<size_t>(cn) * CU) {
          Pack16 p[CU];
          size_t w[CU];
#pragma unroll
          for (int u = 0; u < CU; u++) {
            const size_t j = j0 + static_cast<size_t>(u) * cn;
            w[u] = j < PT_ ? (((j >> tileShift) * G + b) << tileShift) + (j & (T_ - 1)) : chunkVecs;
            const size_t g = g0 + w[u];
            if (w[u] < chunkVecs && g < ngroups) {
              p[u] = loadGroupP(in, g, count, aligned);
              if (extra.n > 0) {
                typename PT::AccPack acc = PT::widen(p[u]);
                for (int k = 0; k < extra.n; k++) {
                  PT::combine(acc, loadGroupP(static_cast<const T*>(extra.p[k]), g, count, aligned), op);
                }
                p[u] = PT::narrow(acc);
              }
            } else {
              p[u].w[0] = p[u].w[1] = p[u].w[2] = p[u].w[3] = 0u;  // keep the reduction of the padding finite
            }
          }
#pragma unroll
          for (int u = 0; u < CU; u++) {
            if (w[u] < chunkVecs) st128(slot + w[u] * 16, p[u]);
          }
        }
      }
    }
    used++;
    if (!blockBarrier(a, e + used)) break;
  }
  retire(a, static_cast<uint32_t>(nchunks + 2), 0);
}

namespace {
template <typename T>
const void* pipeFn(bool mc, int nranks) {
  constexpr bool hot = std::is_same<T, float>::value || std::is_same<T, __half>::value || std::is_same<T, __nv_bfloat16>::value;
  if constexpr (hot) {
    if (mc) return reinterpret_cast<const void*>(pipelinedAllreduceKernel<T, true, 0>);
    if (nranks == 2) return reinterpret_cast<const void*>(pipelinedAllreduceKernel<T, false, 2>);
    if (nranks == 4) return reinterpret_cast<const void*>(pipelinedAllreduceKernel<T, false, 4>);
    if (nranks == 8) return reinterpret_cast<const void*>(pipelinedAllreduceKernel<T, false, 8>);
  }
  return reinterpret_cast<const void*>(pipelinedAllreduceKernel<T, false, 0>);
}

const void* pipeKernelForImpl(DataType dt, bool mc, int nranks) {
  switch (dt) {
    case DataType::INT8: return pipeFn<int8_t>(false, nranks);
    case DataType::UINT8: return pipeFn<uint8_t>(false, nranks);
    case DataType::INT16: return pipeFn<int16_t>(false, nranks);
    case DataType::INT32: return pipeFn<int32_t>(false, nranks);
    case DataType::UINT32: return pipeFn<uint32_t>(false, nranks);
    case DataType::INT64: return pipeFn<long long>(false, nranks);
    case DataType::UINT64: return pipeFn<unsigned long long>(false, nranks);
    case DataType::FLOAT32: return pipeFn<float>(mc, nranks);
    case DataType::FLOAT64: return pipeFn<double>(false, nranks);
    case DataType::FLOAT16: return pipeFn<__half>(mc, nranks);
    case DataType::BFLOAT16: return pipeFn<__nv_bfloat16>(mc, nranks);
  }
  return nullptr;
}
}  // namespace

const void* pipelinedKernelFor(DataType dt, bool mc, int nranks) { return pipeKernelForImpl(dt, mc, nranks); }

void preloadPipelineKernels() {
  for (DataType dt : {DataType::INT8, DataType::UINT8, DataType::INT16, DataType::INT32, DataType::UINT32, DataType::INT64,
                      DataType::UINT64, DataType::FLOAT32, DataType::FLOAT64, DataType::FLOAT16, DataType::BFLOAT16}) {
    for (bool mc : {false, true}) {
      for (int nr : {0, 2, 4, 8}) {
        cudaFuncAttributes attr;
        cudaFuncGetAttributes(&attr, pipeKernelForImpl(dt, mc, nr));
      }
    }
  }
  cudaGetLastError();
}

void launchPipelinedAllreduce(const CommArgs& a, const void* in, void* out, size_t count, DataType dt, ReduceOp op,
                              float scale, const PeerPtrs& stage, void* mcStage, int tileVecs, int exchangeThreads,
                              const LocalPtrs& extra, int blocks, cudaStream_t stream) {
  const bool mc = mcStage != nullptr && nvlsSupports(dt, op);
  const void* k = pipeKernelForImpl(dt, mc, a.nranks);
  CommArgs ca = a;
  PeerPtrs st = stage;
  LocalPtrs ex = extra;
  char* mcp = mc ? static_cast<char*>(mcStage) : nullptr;
  DevOp dop = static_cast<DevOp>(op);
  void* args[] = {&ca, &in, &out, &count, &dop, &scale, &st, &mcp, &tileVecs, &exchangeThreads, &ex};
  cudaLaunchKernel(k, dim3(static_cast<unsigned>(blocks)), dim3(kThreads), args, 0, stream);
}

}  // namespace cuda
}  // namespace glb
