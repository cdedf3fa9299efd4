// Generic step-table executor for the literal allreduce schedules (see schedules.h).
#include "glb/cuda/device_common.cuh"
#include "glb/cuda/schedules.h"

namespace glb {
namespace cuda {

template <typename T>
__global__ void __launch_bounds__(kThreads)
scheduleKernel(CommArgs a, PeerPtrs bufs, PeerPtrs stage, const SchedStep* __restrict__ table, int nsteps, DevOp op,
               bool vectorOk) {
  using PT = PackTraits<T>;
  const uint32_t e = loadEpoch(a);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  T* local = static_cast<T*>(bufs.p[a.rank]);
  T* myStage = static_cast<T*>(stage.p[a.rank]);

  for (int s = 0; s < nsteps; s++) {
    // Step s may read what peers produced in step s-1 (and, for s == 0, their inputs).
    blockBarrier(a, e + 1 + s);
    const SchedStep st = table[s];
    const size_t off = st.off, len = st.len;
    if (len == 0) continue;
    const bool v16 = vectorOk && (off * sizeof(T)) % 16 == 0;
    const size_t nvec = v16 ? len / PT::kElems : 0;
    char* dst = reinterpret_cast<char*>((st.mode == SCHED_STAGE ? myStage : local) + off);
    if (st.mode == SCHED_STAGE) {
      const char* src = reinterpret_cast<const char*>(local + off);
      for (size_t v = tid; v < nvec; v += nthreads) st128(dst + v * 16, ld128_stream(src + v * 16));
      for (size_t i = nvec * PT::kElems + tid; i < len; i += nthreads) myStage[off + i] = local[off + i];
      continue;
    }
    const PeerPtrs& srcs = st.fromStage ? stage : bufs;
    if (st.mode == SCHED_COPY) {
      const char* src = reinterpret_cast<const char*>(static_cast<const T*>(srcs.p[st.peers[0]]) + off);
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
          if (v < nvec) st128(dst + v * 16, p[u]);
        }
      }
      for (size_t i = nvec * PT::kElems + tid; i < len; i += nthreads) {
        local[off + i] = static_cast<const T*>(srcs.p[st.peers[0]])[off + i];
      }
    } else {
      for (size_t v = tid; v < nvec; v += nthreads) {
        typename PT::AccPack acc = PT::widen(ld128(dst + v * 16));
        for (int p = 0; p < st.npeers; p++) {
          PT::combine(acc, ld128_stream(reinterpret_cast<const char*>(static_cast<const T*>(srcs.p[st.peers[p]]) + off) + v * 16), op);
        }
        st128(dst + v * 16, PT::narrow(acc));
      }
      for (size_t i = nvec * PT::kElems + tid; i < len; i += nthreads) {
        T acc = local[off + i];
        for (int p = 0; p < st.npeers; p++) acc = PT::combineOne(acc, static_cast<const T*>(srcs.p[st.peers[p]])[off + i], op);
        local[off + i] = acc;
      }
    }
  }
  blockBarrier(a, e + 1 + nsteps);
  retire(a, nsteps + 1, 0);
}

void launchSchedule(const CommArgs& a, const PeerPtrs& bufs, const PeerPtrs& stage, const SchedStep* table,
                    int nsteps, DataType dt, ReduceOp op, bool vectorOk, int blocks, cudaStream_t stream) {
  const DevOp dop = static_cast<DevOp>(op);
#define GLB_CASE(E, T)                                                                                          \
  case DataType::E:                                                                                             \
    scheduleKernel<T><<<blocks, kThreads, 0, stream>>>(a, bufs, stage, table, nsteps, dop, vectorOk);           \
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

}  // namespace cuda
}  // namespace glb
