// Data-movement collectives over NVLink peer memory (sm_100a): broadcast,
// allgather(v), alltoall(v), reduce_scatter, reduce, gather(v), scatter. All are
// single kernels bracketed by device-side flag barriers; payload moves with
// 128-bit peer loads/stores (or multimem.st through the NVSwitch where a multicast
// alias exists). None of these has a CUDA variant in the reference.
#include "glb/cuda/device_common.cuh"
#include "glb/cuda/kernels.h"

namespace glb {
namespace cuda {

namespace {

__device__ __forceinline__ void shareOfB(size_t n, int parts, int r, size_t& begin, size_t& end) {
  const size_t base = n / parts, rem = n % parts;
  begin = r * base + (static_cast<size_t>(r) < rem ? r : rem);
  end = begin + base + (static_cast<size_t>(r) < rem ? 1 : 0);
}

template <typename W>
__device__ __forceinline__ void gridCopyWords(char* dst, const char* src, size_t bytes, size_t tid, size_t nthreads) {
  const size_t n = bytes / sizeof(W);
  constexpr int U = 4;
  for (size_t i0 = tid; i0 < n; i0 += nthreads * U) {
    W w[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + u * nthreads;
      if (i < n) w[u] = reinterpret_cast<const W*>(src)[i];
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = i0 + u * nthreads;
      if (i < n) reinterpret_cast<W*>(dst)[i] = w[u];
    }
  }
  for (size_t i = n * sizeof(W) + tid; i < bytes; i += nthreads) dst[i] = src[i];
}

// Copy `bytes` from src to dst with the whole grid, using the widest access both
// pointers are aligned for (16 B packs when `vec`, else 8 / 4 / 2 / 1 bytes).
__device__ __forceinline__ void gridCopy(char* dst, const char* src, size_t bytes, bool vec, size_t tid,
                                         size_t nthreads) {
  const uintptr_t both = reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src);
  if (vec && both % 16 == 0) {
    const size_t nvec = bytes / 16;
    constexpr int U = 4;
    for (size_t v0 = tid; v0 < nvec; v0 += nthreads * U) {
      Pack16 p[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = v0 + u * nthreads;
        if (v < nvec) p[u] = ld128_stream(src + v * 16);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = v0 + u * nthreads;
        if (v < nvec) st128_stream(dst + v * 16, p[u]);
      }
    }
    for (size_t i = nvec * 16 + tid; i < bytes; i += nthreads) dst[i] = src[i];
  } else if (both % 8 == 0) {
    gridCopyWords<unsigned long long>(dst, src, bytes, tid, nthreads);
  } else if (both % 4 == 0) {
    gridCopyWords<unsigned int>(dst, src, bytes, tid, nthreads);
  } else if (both % 2 == 0) {
    gridCopyWords<unsigned short>(dst, src, bytes, tid, nthreads);
  } else {
    for (size_t i = tid; i < bytes; i += nthreads) dst[i] = src[i];
  }
}

}  // namespace

// ---- broadcast -----------------------------------------------------------------------
// mode 0: root pushes the whole buffer to every peer (latency regime).
// mode 1: scatter + allgather: root pushes slice i to rank i, then every rank
//         pushes its slice to the others; per-GPU egress ~S instead of (P-1)·S.
// mode 2: NVLS: root issues multimem.st, the switch replicates.
__global__ void __launch_bounds__(kThreads)
broadcastKernel(CommArgs a, PeerPtrs bufs, char* mc, size_t bytes, int root, int mode, bool vec) {
  const uint32_t e = loadEpoch(a);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int P = a.nranks;
  blockBarrier<false>(a, e + 1);  // every destination may now be overwritten
  uint32_t used = 2;
  if (mode == 0) {
    if (a.rank == root) {
      const char* src = static_cast<const char*>(bufs.p[root]);
      const size_t nvec = vec ? bytes / 16 : 0;
      constexpr int U = 4;
      for (size_t v0 = tid; v0 < nvec; v0 += nthreads * U) {
        Pack16 p[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = v0 + u * nthreads;
          if (v < nvec) p[u] = ld128_stream(src + v * 16);
        }
        for (int i = 1; i < P; i++) {
          char* dst = static_cast<char*>(bufs.p[(root + i) % P]);
#pragma unroll
          for (int u = 0; u < U; u++) {
            const size_t v = v0 + u * nthreads;
            if (v < nvec) st128_stream(dst + v * 16, p[u]);
          }
        }
      }
      for (size_t i = nvec * 16 + tid; i < bytes; i += nthreads) {
        const char c = src[i];
        for (int r = 1; r < P; r++) static_cast<char*>(bufs.p[(root + r) % P])[i] = c;
      }
    }
  } else if (mode == 2) {
    if (a.rank == root) {
      const char* src = static_cast<const char*>(bufs.p[root]);
      const size_t nvec = bytes / 16;
      constexpr int U = 4;
      for (size_t v0 = tid; v0 < nvec; v0 += nthreads * U) {
        Pack16 p[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = v0 + u * nthreads;
          if (v < nvec) p[u] = ld128_stream(src + v * 16);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = v0 + u * nthreads;
          if (v < nvec) multimemSt128(mc + v * 16, p[u]);
        }
      }
      for (size_t i = nvec * 16 + tid; i < bytes; i += nthreads) {
        const char c = src[i];
        for (int r = 1; r < P; r++) static_cast<char*>(bufs.p[(root + r) % P])[i] = c;
      }
    }
  } else {
    // Slices are 16-byte granular so every phase stays vectorised.
    const size_t units = bytes / 16;
    if (a.rank == root) {
      for (int i = 1; i < P; i++) {
        const int dst = (root + i) % P;
        size_t b, en;
        shareOfB(units, P, dst, b, en);
        gridCopy(static_cast<char*>(bufs.p[dst]) + b * 16, static_cast<const char*>(bufs.p[root]) + b * 16,
                 (en - b) * 16, vec, tid, nthreads);
      }
      // Unaligned tail goes directly.
      for (size_t i = units * 16 + tid; i < bytes; i += nthreads) {
        const char c = static_cast<const char*>(bufs.p[root])[i];
        for (int r = 1; r < P; r++) static_cast<char*>(bufs.p[(root + r) % P])[i] = c;
      }
    }
    blockBarrier(a, e + 2);
    used = 3;
    // Every rank (the root too: nobody else holds its slice) now pushes the slice it
    // owns to all ranks that still miss it, i.e. everyone but itself and the root.
    {
      size_t b, en;
      shareOfB(units, P, a.rank, b, en);
      const char* src = static_cast<const char*>(bufs.p[a.rank]) + b * 16;
      const size_t nvec = vec ? (en - b) : 0;
      for (size_t v = tid; v < nvec; v += nthreads) {
        const Pack16 p = ld128_stream(src + v * 16);
        for (int i = 1; i < P; i++) {
          const int dst = (a.rank + i) % P;
          if (dst != root) st128_stream(static_cast<char*>(bufs.p[dst]) + b * 16 + v * 16, p);
        }
      }
      for (size_t i = nvec * 16 + tid; i < (en - b) * 16; i += nthreads) {
        const char c = src[i];
        for (int r = 1; r < P; r++) {
          const int dst = (a.rank + r) % P;
          if (dst != root) (static_cast<char*>(bufs.p[dst]) + b * 16)[i] = c;
        }
      }
    }
  }
  blockBarrier(a, e + used);
  retire(a, used, 0);
}

void launchBroadcast(const CommArgs& a, const PeerPtrs& bufs, void* mc, size_t bytes, int root, int mode, bool vec,
                     int blocks, cudaStream_t stream) {
  broadcastKernel<<<blocks, kThreads, 0, stream>>>(a, bufs, static_cast<char*>(mc), bytes, root, mode, vec);
}

// ---- allgather(v) / gather(v): push my block into peers' outputs ---------------------------
// Rank r's input block (inBytes[r] bytes at `in`) lands at outOff[r] in the output
// of every rank in [dstBegin, dstEnd) — all ranks for allgather, just the root for
// gather. Outputs are peer-mapped; the input is only read locally.
struct VArgs {
  size_t off[kMaxRanks];
  size_t len[kMaxRanks];
};

__global__ void __launch_bounds__(kThreads)
gatherPushKernel(CommArgs a, const char* __restrict__ in, PeerPtrs outs, char* mcOut, VArgs va, int onlyDst,
                 bool vec) {
  const uint32_t e = loadEpoch(a);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int P = a.nranks;
  blockBarrier<false>(a, e + 1);
  const size_t off = va.off[a.rank];
  const size_t len = va.len[a.rank];
  const bool v16 = vec && off % 16 == 0;
  const size_t nvec = v16 ? len / 16 : 0;
  if (onlyDst >= 0) {
    gridCopy(static_cast<char*>(outs.p[onlyDst]) + off, in, len, v16, tid, nthreads);
  } else if (mcOut != nullptr && v16) {
    for (size_t v = tid; v < nvec; v += nthreads) multimemSt128(mcOut + off + v * 16, ld128_stream(in + v * 16));
    for (size_t i = nvec * 16 + tid; i < len; i += nthreads) {
      const char c = in[i];
      for (int r = 0; r < P; r++) (static_cast<char*>(outs.p[r]) + off)[i] = c;
    }
  } else {
    constexpr int U = 4;
    for (size_t v0 = tid; v0 < nvec; v0 += nthreads * U) {
      Pack16 p[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t v = v0 + u * nthreads;
        if (v < nvec) p[u] = ld128_stream(in + v * 16);
      }
      for (int i = 0; i < P; i++) {
        char* dst = static_cast<char*>(outs.p[(a.rank + i) % P]) + off;
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = v0 + u * nthreads;
          if (v < nvec) st128_stream(dst + v * 16, p[u]);
        }
      }
    }
    for (size_t i = nvec * 16 + tid; i < len; i += nthreads) {
      const char c = in[i];
      for (int r = 0; r < P; r++) (static_cast<char*>(outs.p[r]) + off)[i] = c;
    }
  }
  blockBarrier(a, e + 2);
  retire(a, 2, 0);
}

void launchGatherPush(const CommArgs& a, const void* in, const PeerPtrs& outs, void* mcOut, const size_t* offs,
                      const size_t* lens, int onlyDst, bool vec, int blocks, cudaStream_t stream) {
  VArgs va;
  for (int i = 0; i < kMaxRanks; i++) {
    va.off[i] = i < a.nranks ? offs[i] : 0;
    va.len[i] = i < a.nranks ? lens[i] : 0;
  }
  gatherPushKernel<<<blocks, kThreads, 0, stream>>>(a, static_cast<const char*>(in), outs, static_cast<char*>(mcOut),
                                                    va, onlyDst, vec);
}

// ---- alltoall(v) / scatter: push chunk j of my input into slot (me) of rank j's output -------
// sendOff/sendLen index my input per destination; recvOff[r] (same on all ranks by
// symmetry of the exchange tables) is where MY chunk lands in rank r's output.
__global__ void __launch_bounds__(kThreads)
alltoallPushKernel(CommArgs a, const char* __restrict__ in, PeerPtrs outs, VArgs send, VArgs dstOff, VArgs myRecv,
                   VArgs blk, bool exchange, int onlySrc, bool vec) {
  const uint32_t e = loadEpoch(a);
  const int P = a.nranks;
  // v-variant: publish where I expect each source's chunk; peers read it after the
  // barrier (every CTA writes the same values, so no cross-CTA ordering is needed).
  if (exchange && threadIdx.x < P) a.sig[a.rank]->xchg[threadIdx.x] = myRecv.off[threadIdx.x];
  blockBarrier(a, e + 1);
  if (onlySrc < 0 || onlySrc == a.rank) {
    // CTAs are dealt to destinations in proportion to the bytes each one gets (the host
    // fills blk.off = first CTA, blk.len = CTA count per destination), so every link is
    // busy at once and uneven (v-variant) splits stay balanced. With fewer CTAs than
    // destinations every CTA walks all destinations.
    const bool partitioned = blk.len[0] + blk.off[P - 1] > 0 && static_cast<size_t>(gridDim.x) >= static_cast<size_t>(P);
    for (int i = 0; i < P; i++) {
      const int dst = (a.rank + i) % P;
      size_t tid, nthreads;
      if (partitioned) {
        const size_t b0 = blk.off[dst], nb = blk.len[dst];
        if (blockIdx.x < b0 || blockIdx.x >= b0 + nb) continue;
        tid = (blockIdx.x - b0) * blockDim.x + threadIdx.x;
        nthreads = nb * blockDim.x;
      } else {
        tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
        nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
      }
      const size_t so = send.off[dst], sl = send.len[dst];
      const size_t dofs = exchange ? static_cast<size_t>(*reinterpret_cast<volatile unsigned long long*>(
                                         &a.sig[dst]->xchg[a.rank]))
                                   : dstOff.off[dst];
      gridCopy(static_cast<char*>(outs.p[dst]) + dofs, in + so, sl, vec, tid, nthreads);
    }
  }
  blockBarrier(a, e + 2);
  retire(a, 2, 0);
}

void launchAlltoallPush(const CommArgs& a, const void* in, const PeerPtrs& outs, const size_t* sendOff,
                        const size_t* sendLen, const size_t* dstOff, const size_t* recvOffTable, int onlySrc,
                        bool vec, int blocks, cudaStream_t stream) {
  VArgs s, d, rcv;
  for (int i = 0; i < kMaxRanks; i++) {
    s.off[i] = i < a.nranks ? sendOff[i] : 0;
    s.len[i] = i < a.nranks ? sendLen[i] : 0;
    d.off[i] = i < a.nranks ? dstOff[i] : 0;
    d.len[i] = 0;
    rcv.off[i] = (recvOffTable != nullptr && i < a.nranks) ? recvOffTable[i] : 0;
    rcv.len[i] = 0;
  }
  // CTA ranges per destination, proportional to payload (at least one each).
  VArgs blk;
  for (int i = 0; i < kMaxRanks; i++) blk.off[i] = blk.len[i] = 0;
  if (blocks >= a.nranks) {
    size_t total = 0;
    for (int i = 0; i < a.nranks; i++) total += s.len[i];
    size_t used = 0;
    int spare = blocks - a.nranks;
    for (int i = 0; i < a.nranks; i++) {
      size_t extra = total > 0 ? static_cast<size_t>(spare) * s.len[i] / total : 0;
      blk.off[i] = used;
      blk.len[i] = 1 + extra;
      used += blk.len[i];
    }
  }
  alltoallPushKernel<<<blocks, kThreads, 0, stream>>>(a, static_cast<const char*>(in), outs, s, d, rcv, blk,
                                                      recvOffTable != nullptr, onlySrc, vec);
}

// ---- reduce_scatter / reduce: pull my slice from every peer's input, reduce, store -----------
// Inputs are peer-mapped. Rank r reduces elements [elemOff[r], elemOff[r]+elemLen[r])
// and writes them to `out` (local pointer; for reduce() it is the root's peer-mapped
// output at the same offset).
struct EArgs {
  size_t off[kMaxRanks];
  size_t len[kMaxRanks];
};

template <typename T, int NR, int UNROLL>
__global__ void __launch_bounds__(kThreads)
reducePullKernel(CommArgs a, PeerPtrs ins, char* mcIn, T* __restrict__ out, EArgs ea, DevOp op, bool vec,
                 bool useMc) {
  using PT = PackTraits<T>;
  const uint32_t e = loadEpoch(a);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  const int P = NR > 0 ? NR : a.nranks;
  blockBarrier<false>(a, e + 1);
  const size_t off = ea.off[a.rank], len = ea.len[a.rank];
  const bool v16 = vec && (off * sizeof(T)) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
  const size_t nvec = v16 ? len / PT::kElems : 0;
  const size_t byteOff = off * sizeof(T);
  if (useMc && v16) {
    if constexpr (std::is_same<T, float>::value || std::is_same<T, __half>::value ||
                  std::is_same<T, __nv_bfloat16>::value) {
      constexpr int U = 4;
      for (size_t v0 = tid; v0 < nvec; v0 += nthreads * U) {
        Pack16 r[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = v0 + u * nthreads;
          if (v < nvec) r[u] = Multimem<T>::ldReduceAdd(mcIn + byteOff + v * 16);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t v = v0 + u * nthreads;
          if (v < nvec) st128(reinterpret_cast<char*>(out) + v * 16, r[u]);
        }
      }
    }
  } else {
    constexpr int kSlots = NR > 0 ? NR : kMaxRanks;
    const char* peer[kSlots];
#pragma unroll
    for (int i = 0; i < kSlots; i++) {
      peer[i] = i < P ? static_cast<const char*>(ins.p[(a.rank + i) % P]) + byteOff : nullptr;
    }
    for (size_t v0 = tid; v0 < nvec; v0 += nthreads * UNROLL) {
      Pack16 p[UNROLL][kSlots];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const size_t v = v0 + static_cast<size_t>(u) * nthreads;
        if (v < nvec) {
#pragma unroll
          for (int i = 0; i < kSlots; i++) {
            if (i < P) p[u][i] = ld128_stream(peer[i] + v * 16);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        const size_t v = v0 + static_cast<size_t>(u) * nthreads;
        if (v < nvec) {
          typename PT::AccPack acc = PT::widen(p[u][0]);
#pragma unroll
          for (int i = 1; i < kSlots; i++) {
            if (i < P) PT::combine(acc, p[u][i], op);
          }
          st128(reinterpret_cast<char*>(out) + v * 16, PT::narrow(acc));
        }
      }
    }
  }
  for (size_t i = nvec * PT::kElems + tid; i < len; i += nthreads) {
    T acc = static_cast<const T*>(ins.p[0])[off + i];
    for (int r = 1; r < P; r++) acc = PT::combineOne(acc, static_cast<const T*>(ins.p[r])[off + i], op);
    out[i] = acc;
  }
  blockBarrier(a, e + 2);
  retire(a, 2, 0);
}

namespace {
template <typename T>
void launchReducePullT(const CommArgs& a, const PeerPtrs& ins, char* mc, void* out, const EArgs& ea, DevOp dop, bool vec,
                       bool useMc, int blocks, cudaStream_t stream) {
  constexpr bool hot = std::is_same<T, float>::value || std::is_same<T, __half>::value ||
                       std::is_same<T, __nv_bfloat16>::value;
  if constexpr (hot) {
    switch (a.nranks) {
      case 2: reducePullKernel<T, 2, 4><<<blocks, kThreads, 0, stream>>>(a, ins, mc, static_cast<T*>(out), ea, dop, vec, useMc); return;
      case 4: reducePullKernel<T, 4, 2><<<blocks, kThreads, 0, stream>>>(a, ins, mc, static_cast<T*>(out), ea, dop, vec, useMc); return;
      case 8: reducePullKernel<T, 8, 2><<<blocks, kThreads, 0, stream>>>(a, ins, mc, static_cast<T*>(out), ea, dop, vec, useMc); return;
      default: break;
    }
  }
  reducePullKernel<T, 0, 1><<<blocks, kThreads, 0, stream>>>(a, ins, mc, static_cast<T*>(out), ea, dop, vec, useMc);
}
}  // namespace

void launchReducePull(const CommArgs& a, const PeerPtrs& ins, void* mcIn, void* out, const size_t* elemOff,
                      const size_t* elemLen, DataType dt, ReduceOp op, bool vec, bool useMc, int blocks,
                      cudaStream_t stream) {
  EArgs ea;
  for (int i = 0; i < kMaxRanks; i++) {
    ea.off[i] = i < a.nranks ? elemOff[i] : 0;
    ea.len[i] = i < a.nranks ? elemLen[i] : 0;
  }
  const DevOp dop = static_cast<DevOp>(op);
  char* mc = static_cast<char*>(mcIn);
#define GLB_CASE(E, T)                                                              \
  case DataType::E:                                                                 \
    launchReducePullT<T>(a, ins, mc, out, ea, dop, vec, useMc, blocks, stream);     \
    break;
  switch (dt) {
    GLB_CASE(INT8, int8_t)
    GLB_CASE(UINT8, uint8_t)
    GLB_CASE(INT16, int16_t)
    GLB_CASE(INT32, int32_t)
    GLB_CASE(UINT32, uint32_t)
    GLB_CASE(INT64, long long)
    GLB_CASE(UINT64, unsigned long long)
    GLB_CASE(FLOAT32, float)
    GLB_CASE(FLOAT64, double)
    GLB_CASE(FLOAT16, __half)
    GLB_CASE(BFLOAT16, __nv_bfloat16)
  }
#undef GLB_CASE
}


void preloadCollectiveKernels() {
  auto touch = [](const void* k) {
    cudaFuncAttributes attr;
    cudaFuncGetAttributes(&attr, k);
  };
  touch(reinterpret_cast<const void*>(broadcastKernel));
  touch(reinterpret_cast<const void*>(gatherPushKernel));
  touch(reinterpret_cast<const void*>(alltoallPushKernel));
#define GLB_TOUCH_RP(T) touch(reinterpret_cast<const void*>(reducePullKernel<T, 0, 1>));
  GLB_TOUCH_RP(int8_t) GLB_TOUCH_RP(uint8_t) GLB_TOUCH_RP(int16_t) GLB_TOUCH_RP(int32_t) GLB_TOUCH_RP(uint32_t)
  GLB_TOUCH_RP(long long) GLB_TOUCH_RP(unsigned long long) GLB_TOUCH_RP(float) GLB_TOUCH_RP(double)
  GLB_TOUCH_RP(__half) GLB_TOUCH_RP(__nv_bfloat16)
#undef GLB_TOUCH_RP
#define GLB_TOUCH_HOT(T)                                                   \
  touch(reinterpret_cast<const void*>(reducePullKernel<T, 2, 4>));         \
  touch(reinterpret_cast<const void*>(reducePullKernel<T, 4, 2>));         \
  touch(reinterpret_cast<const void*>(reducePullKernel<T, 8, 2>));
  GLB_TOUCH_HOT(float) GLB_TOUCH_HOT(__half) GLB_TOUCH_HOT(__nv_bfloat16)
#undef GLB_TOUCH_HOT
  cudaGetLastError();
}

}  // namespace cuda
}  // namespace glb
