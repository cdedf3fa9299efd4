// Device-side building blocks shared by every collective kernel (sm_100a):
//
//  * CommArgs / SignalPad — per-rank signal page mapped into every peer. A kernel
//    synchronises with its peers by storing a monotonically increasing epoch into
//    the peers' pads (st.release.sys) and spinning on its own pad
//    (ld.acquire.sys): this replaces the host-side stream.wait() + notification
//    Buffers of the reference (cuda_allreduce_ring_chunked.cc:177-208).
//  * 16-byte vector packs with fp32 accumulation for 16-bit types and a runtime
//    reduction op (uniform branch; these kernels are link/HBM bound).
//  * multimem.* wrappers for the NVLS (NVSwitch multicast) path.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "glb/cuda/comm_types.h"
#include "glb/types.h"

namespace glb {
namespace cuda {

enum class DevOp : int { SUM = 1, PRODUCT = 2, MAX = 3, MIN = 4 };

// ---- memory-ordering primitives -------------------------------------------------

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs with the same blockIdx on every rank rendezvous. Everything the
// callers wrote before (including stores into peer memory) is visible to every
// peer's block after it returns (bar.sync + cumulative release/acquire at .sys).
// kRelease=false is for the FIRST barrier of a kernel that has not written anything a
// peer will read: "my kernel has started" needs no fence (the inputs were produced by
// earlier kernels and are already visible system-wide), which saves a MEMBAR.SYS.
template <bool kRelease = true>
__device__ __forceinline__ void blockBarrier(const CommArgs& a, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < a.nranks) {
    const int peer = threadIdx.x;
    if (kRelease) {
      st_release_sys(&a.sig[peer]->flag[blockIdx.x][a.rank], epoch);
    } else {
      st_relaxed_sys(&a.sig[peer]->flag[blockIdx.x][a.rank], epoch);
    }
    const uint32_t* mine = &a.sig[a.rank]->flag[blockIdx.x][peer];
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
    }
  }
  __syncthreads();
}

// Read the epoch at kernel entry (all CTAs see the same value: it only changes
// when the LAST CTA of a launch retires, after every CTA has read it).
__device__ __forceinline__ uint32_t loadEpoch(const CommArgs& a) { return ld_relaxed_sys(&a.sig[a.rank]->epoch); }

// Called by every CTA at the very end; the last one publishes the new counters.
__device__ __forceinline__ void retire(const CommArgs& a, uint32_t barriersUsed, uint32_t stagedLaunch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    SignalPad* me = a.sig[a.rank];
    // No fences: the ticket is an L2 atomic (the last CTA sees every earlier arrival) and
    // the counters are next read by the NEXT launch, after this kernel has retired.
    uint32_t ticket = atomicAdd(&me->done, 1u);
    if (ticket == gridDim.x - 1) {
      me->done = 0;
      me->epoch += barriersUsed;
      me->stageSeq += stagedLaunch;
    }
  }
}

// ---- 16-byte packs ----------------------------------------------------------------

struct alignas(16) Pack16 {
  uint32_t w[4];
};

__device__ __forceinline__ Pack16 ld128(const void* p) {
  Pack16 v;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(p)
               : "memory");
  return v;
}
// Streaming variants: do not allocate in L1 (peer / single-use data).
__device__ __forceinline__ Pack16 ld128_stream(const void* p) {
  Pack16 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st128(void* p, const Pack16& v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3])
               : "memory");
}
__device__ __forceinline__ void st128_stream(void* p, const Pack16& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]),
               "r"(v.w[2]), "r"(v.w[3])
               : "memory");
}

template <typename T>
__device__ __forceinline__ T applyOp(T a, T b, DevOp op) {
  switch (op) {
    case DevOp::SUM: return a + b;
    case DevOp::PRODUCT: return a * b;
    case DevOp::MAX: return b > a ? b : a;
    default: return b < a ? b : a;
  }
}

// Accumulator traits: how one 16-byte pack of T is widened, combined, narrowed.
template <typename T>
struct PackTraits {
  static constexpr int kElems = 16 / sizeof(T);
  using Acc = T;
  struct AccPack {
    T v[kElems];
  };
  __device__ static AccPack widen(const Pack16& p) {
    AccPack a;
    const T* t = reinterpret_cast<const T*>(&p);
#pragma unroll
    for (int i = 0; i < kElems; i++) a.v[i] = t[i];
    return a;
  }
  __device__ static void combine(AccPack& a, const Pack16& p, DevOp op) {
    const T* t = reinterpret_cast<const T*>(&p);
#pragma unroll
    for (int i = 0; i < kElems; i++) a.v[i] = applyOp<T>(a.v[i], t[i], op);
  }
  __device__ static Pack16 narrow(const AccPack& a) {
    Pack16 p;
    T* t = reinterpret_cast<T*>(&p);
#pragma unroll
    for (int i = 0; i < kElems; i++) t[i] = a.v[i];
    return p;
  }
  __device__ static T one(const T* p) { return *p; }
  __device__ static T combineOne(T a, T b, DevOp op) { return applyOp<T>(a, b, op); }
};

template <>
struct PackTraits<__half> {
  static constexpr int kElems = 8;
  struct AccPack {
    float v[8];
  };
  __device__ static AccPack widen(const Pack16& p) {
    AccPack a;
    const __half2* h = reinterpret_cast<const __half2*>(&p);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 f = __half22float2(h[i]);
      a.v[2 * i] = f.x;
      a.v[2 * i + 1] = f.y;
    }
    return a;
  }
  __device__ static void combine(AccPack& a, const Pack16& p, DevOp op) {
    const __half2* h = reinterpret_cast<const __half2*>(&p);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 f = __half22float2(h[i]);
      a.v[2 * i] = applyOp<float>(a.v[2 * i], f.x, op);
      a.v[2 * i + 1] = applyOp<float>(a.v[2 * i + 1], f.y, op);
    }
  }
  __device__ static Pack16 narrow(const AccPack& a) {
    Pack16 p;
    __half2* h = reinterpret_cast<__half2*>(&p);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(a.v[2 * i], a.v[2 * i + 1]);
    return p;
  }
  __device__ static __half combineOne(__half a, __half b, DevOp op) {
    return __float2half_rn(applyOp<float>(__half2float(a), __half2float(b), op));
  }
};

template <>
struct PackTraits<__nv_bfloat16> {
  static constexpr int kElems = 8;
  struct AccPack {
    float v[8];
  };
  __device__ static AccPack widen(const Pack16& p) {
    AccPack a;
    // bf16 -> fp32 is a 16-bit shift.
#pragma unroll
    for (int i = 0; i < 4; i++) {
      a.v[2 * i] = __uint_as_float(p.w[i] << 16);
      a.v[2 * i + 1] = __uint_as_float(p.w[i] & 0xffff0000u);
    }
    return a;
  }
  __device__ static void combine(AccPack& a, const Pack16& p, DevOp op) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      a.v[2 * i] = applyOp<float>(a.v[2 * i], __uint_as_float(p.w[i] << 16), op);
      a.v[2 * i + 1] = applyOp<float>(a.v[2 * i + 1], __uint_as_float(p.w[i] & 0xffff0000u), op);
    }
  }
  __device__ static Pack16 narrow(const AccPack& a) {
    Pack16 p;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&p);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
    return p;
  }
  __device__ static __nv_bfloat16 combineOne(__nv_bfloat16 a, __nv_bfloat16 b, DevOp op) {
    return __float2bfloat16_rn(applyOp<float>(__bfloat162float(a), __bfloat162float(b), op));
  }
};

// ---- NVLS (multimem) --------------------------------------------------------------
// ld_reduce returns the reduction of the same address across every GPU bound to
// the multicast object (performed inside the NVSwitch); st broadcasts to all.

template <typename T>
struct Multimem;  // only add for float / half / bf16; min/max for half/bf16

template <>
struct Multimem<float> {
  __device__ static Pack16 ldReduceAdd(const void* mc) {
    Pack16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                 : "l"(mc)
                 : "memory");
    return v;
  }
};
template <>
struct Multimem<__half> {
  __device__ static Pack16 ldReduceAdd(const void* mc) {
    Pack16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                 : "l"(mc)
                 : "memory");
    return v;
  }
};
template <>
struct Multimem<__nv_bfloat16> {
  __device__ static Pack16 ldReduceAdd(const void* mc) {
    Pack16 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                 : "l"(mc)
                 : "memory");
    return v;
  }
};

__device__ __forceinline__ void multimemSt128(void* mc, const Pack16& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.w[0]), "r"(v.w[1]),
               "r"(v.w[2]), "r"(v.w[3])
               : "memory");
}

// Map the public dtype tags to device element types.
template <DataType D>
struct DevType;
template <> struct DevType<DataType::INT8> { using type = int8_t; };
template <> struct DevType<DataType::UINT8> { using type = uint8_t; };
template <> struct DevType<DataType::INT16> { using type = int16_t; };
template <> struct DevType<DataType::INT32> { using type = int32_t; };
template <> struct DevType<DataType::UINT32> { using type = uint32_t; };
template <> struct DevType<DataType::INT64> { using type = long long; };
template <> struct DevType<DataType::UINT64> { using type = unsigned long long; };
template <> struct DevType<DataType::FLOAT32> { using type = float; };
template <> struct DevType<DataType::FLOAT64> { using type = double; };
template <> struct DevType<DataType::FLOAT16> { using type = __half; };
template <> struct DevType<DataType::BFLOAT16> { using type = __nv_bfloat16; };

}  // namespace cuda
}  // namespace glb
