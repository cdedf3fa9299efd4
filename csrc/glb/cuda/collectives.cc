#include "glb/cuda/collectives.h"

#include <algorithm>
#include <atomic>

#include "glb/common/utils.h"
#include "glb/cuda/kernels.h"
#include "glb/cuda/trace.h"

namespace glb {
namespace cuda {

const char* allreduceAlgoName(AllreduceAlgo a) {
  switch (a) {
    case AllreduceAlgo::AUTO: return "auto";
    case AllreduceAlgo::ONE_SHOT: return "one_shot";
    case AllreduceAlgo::TWO_SHOT: return "two_shot";
    case AllreduceAlgo::NVLS: return "nvls";
    case AllreduceAlgo::RING: return "ring";
    case AllreduceAlgo::RING_CHUNKED: return "ring_chunked";
    case AllreduceAlgo::HALVING_DOUBLING: return "halving_doubling";
    case AllreduceAlgo::BCUBE: return "bcube";
  }
  return "?";
}

Tuning& tuning() {
  static Tuning t = [] {
    Tuning x;
    x.oneShotMaxBytes = static_cast<size_t>(envInt("CUDA_ONESHOT_MAX", static_cast<long>(x.oneShotMaxBytes)));
    x.nvlsMinBytes = static_cast<size_t>(envInt("CUDA_NVLS_MIN", static_cast<long>(x.nvlsMinBytes)));
    x.maxBlocks = static_cast<int>(envInt("CUDA_BLOCKS", x.maxBlocks));
    x.oneShotBlocks = static_cast<int>(envInt("CUDA_ONESHOT_BLOCKS", x.oneShotBlocks));
    x.copyBlocks = static_cast<int>(envInt("CUDA_COPY_BLOCKS", x.copyBlocks));
    x.nvlsReduceScatter = envFlag("CUDA_NVLS_REDUCE_SCATTER", x.nvlsReduceScatter);
    setOneShotPush(envFlag("CUDA_ONESHOT_PUSH", true));
    x.bcastDirectMaxBytes = static_cast<size_t>(envInt("CUDA_BCAST_DIRECT_MAX", static_cast<long>(x.bcastDirectMaxBytes)));
    return x;
  }();
  return t;
}

namespace {

// Layout of the staging area: two one-shot halves, then the bulk region.
struct StageLayout {
  size_t half;       // bytes of one one-shot half
  size_t bulkOff;    // start of the bulk (two-shot / NVLS staging) region
  size_t bulkBytes;
};

StageLayout layoutOf(const PeerContext& pc) {
  StageLayout l;
  l.half = std::min<size_t>(roundUp(std::max<size_t>(tuning().oneShotMaxBytes, 4096), 4096), pc.stageBytes() / 4);
  l.bulkOff = 2 * l.half;
  l.bulkBytes = pc.stageBytes() - l.bulkOff;
  return l;
}

int blocksFor(const PeerContext& pc, size_t vecsPerRank, int unrollHint, int cap) {
  size_t want = ceilDiv(std::max<size_t>(vecsPerRank, 1), static_cast<size_t>(kThreads) * unrollHint);
  int b = static_cast<int>(std::min<size_t>(want, static_cast<size_t>(cap)));
  return std::max(1, std::min(b, pc.maxBlocks()));
}

std::atomic<uint64_t> gLaunches{0};

void checkLaunch(const char* what) {
  gLaunches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) GLB_THROW(Exception, what, ": kernel launch failed: ", cudaGetErrorString(e));
}

}  // namespace

uint64_t launchCount() { return gLaunches.load(); }
void noteLaunch(unsigned n) { gLaunches.fetch_add(n, std::memory_order_relaxed); }

AllreduceAlgo chooseAllreduce(const PeerContext& pc, size_t bytes, DataType dt, ReduceOp op, bool registered,
                              bool hasMulticast) {
  const auto& t = tuning();
  // One-shot moves (P-1)*S into every GPU, so its break-even shrinks with P. Measured
  // (profiles/runs/sweep_allreduce_{2,8}gpu_f32.json): ~128 KiB at P=2, ~32 KiB at P=8,
  // i.e. oneShotMaxBytes (256 KiB) / P.
  if (bytes <= t.oneShotMaxBytes / static_cast<size_t>(std::max(pc.size, 1)) && bytes <= layoutOf(pc).half) {
    return AllreduceAlgo::ONE_SHOT;
  }
  // In-switch reduction moves ~S(1+1/P) per direction against 2S(P-1)/P for two-shot:
  // a win from P = 4 up, a loss at P = 2 (measured: 994 us vs 645 us for 400 MB).
  if (hasMulticast && pc.size > 2 && bytes >= t.nvlsMinBytes && nvlsSupports(dt, op)) return AllreduceAlgo::NVLS;
  (void)registered;
  return AllreduceAlgo::TWO_SHOT;
}

void barrier(PeerContext& pc, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::barrier");
  DeviceGuard g(pc.device);
  pc.launchGuard();
  launchBarrier(pc.comm(), stream);
  checkLaunch("barrier");
}

void allreduce(PeerContext& pc, const PeerBuffer& buf, size_t byteOffset, size_t count, DataType dt, ReduceOp op,
               AllreduceAlgo algo, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::allreduce");
  if (count == 0) return;
  const size_t es = elementSize(dt);
  const size_t bytes = count * es;
  GLB_ENFORCE_LE(byteOffset + bytes, buf.bytes, "allreduce range exceeds the registered buffer");
  GLB_ENFORCE(op != ReduceOp::CUSTOM, "custom reductions run on the host path only");
  DeviceGuard g(pc.device);
  char* local = static_cast<char*>(buf.local) + byteOffset;
  // Single rank: the buffer already holds the result. GLB_CUDA_FORCE_KERNELS=1 still
  // launches the kernel (loopback on local memory) so it can be profiled with Nsight
  // Compute, which cannot replay kernels that wait for a peer.
  if (pc.size == 1 && !envFlag("CUDA_FORCE_KERNELS", false)) return;
  const bool hasMc = buf.mc != nullptr && byteOffset % 16 == 0;
  if (algo == AllreduceAlgo::AUTO) algo = chooseAllreduce(pc, bytes, dt, op, true, hasMc);
  pc.launchGuard();
  switch (algo) {
    case AllreduceAlgo::ONE_SHOT: {
      const auto l = layoutOf(pc);
      GLB_ENFORCE_LE(bytes, l.half, "one-shot allreduce limited to ", l.half, " bytes");
      const int blocks = blocksFor(pc, bytes / 16, 1, tuning().oneShotBlocks);
      launchOneShotAllreduce(pc.comm(), local, local, count, dt, op, pc.stagePtrs(0), l.half, blocks, stream);
      break;
    }
    case AllreduceAlgo::NVLS: {
      GLB_ENFORCE(hasMc, "NVLS allreduce needs a multicast-bound (symmetric) buffer");
      GLB_ENFORCE(nvlsSupports(dt, op), "NVLS supports sum over float32/float16/bfloat16 only");
      const int blocks = blocksFor(pc, ceilDiv(bytes, 16) / pc.size, 4, tuning().maxBlocks);
      launchNvlsAllreduce(pc.comm(), static_cast<char*>(buf.mc) + byteOffset, buf.ptrsAt(byteOffset), count, dt, blocks,
                          stream);
      break;
    }
    case AllreduceAlgo::TWO_SHOT: {
      const bool vectorOk = buf.vectorOk && byteOffset % 16 == 0;
      const int blocks = blocksFor(pc, bytes / 16 / pc.size, 2, tuning().maxBlocks);
      launchTwoShotAllreduce(pc.comm(), buf.ptrsAt(byteOffset), count, dt, op, vectorOk, blocks, stream);
      break;
    }
    default:
      GLB_THROW_INVALID_OPERATION_EXCEPTION("allreduce: algorithm ", allreduceAlgoName(algo),
                                            " is not available on this entry point");
  }
  checkLaunch("allreduce");
}

void allreduce(PeerContext& pc, const void* in, void* out, size_t count, DataType dt, ReduceOp op,
               AllreduceAlgo algo, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::allreduce(staged)");
  if (count == 0) return;
  GLB_ENFORCE(op != ReduceOp::CUSTOM, "custom reductions run on the host path only");
  DeviceGuard g(pc.device);
  const size_t es = elementSize(dt);
  const size_t bytes = count * es;
  if (pc.size == 1) {
    if (in != out) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  const auto l = layoutOf(pc);
  const bool mcOk = pc.nvlsAvailable();
  if (algo == AllreduceAlgo::AUTO) algo = chooseAllreduce(pc, bytes, dt, op, false, mcOk);
  if (algo == AllreduceAlgo::ONE_SHOT && bytes <= l.half) {
    pc.launchGuard();
    const int blocks = blocksFor(pc, bytes / 16, 1, tuning().oneShotBlocks);
    launchOneShotAllreduce(pc.comm(), in, out, count, dt, op, pc.stagePtrs(0), l.half, blocks, stream);
    checkLaunch("allreduce(one-shot)");
    return;
  }
  if (algo == AllreduceAlgo::ONE_SHOT) algo = AllreduceAlgo::TWO_SHOT;
  GLB_ENFORCE(algo == AllreduceAlgo::TWO_SHOT || algo == AllreduceAlgo::NVLS,
              "allreduce: algorithm ", allreduceAlgoName(algo), " needs registered buffers");
  if (algo == AllreduceAlgo::NVLS) {
    GLB_ENFORCE(mcOk, "NVLS not available on this context");
    GLB_ENFORCE(nvlsSupports(dt, op), "NVLS supports sum over float32/float16/bfloat16 only");
  }
  // Bulk path: copy-in -> fused kernel on the pool -> copy-out, one piece at a time.
  const size_t pieceElems = std::max<size_t>(1, (l.bulkBytes / 16 * 16) / es);
  PeerPtrs stage = pc.stagePtrs(l.bulkOff);
  char* myStage = static_cast<char*>(stage.p[pc.rank]);
  for (size_t done = 0; done < count; done += pieceElems) {
    const size_t n = std::min(pieceElems, count - done);
    const char* src = static_cast<const char*>(in) + done * es;
    char* dst = static_cast<char*>(out) + done * es;
    GLB_CUDA_CHECK(cudaMemcpyAsync(myStage, src, n * es, cudaMemcpyDeviceToDevice, stream));
    pc.launchGuard();
    if (algo == AllreduceAlgo::NVLS) {
      const int blocks = blocksFor(pc, ceilDiv(n * es, 16) / pc.size, 4, tuning().maxBlocks);
      launchNvlsAllreduce(pc.comm(), pc.stageMc(l.bulkOff), stage, n, dt, blocks, stream);
    } else {
      const int blocks = blocksFor(pc, n * es / 16 / pc.size, 2, tuning().maxBlocks);
      launchTwoShotAllreduce(pc.comm(), stage, n, dt, op, true, blocks, stream);
    }
    checkLaunch("allreduce(staged)");
    GLB_CUDA_CHECK(cudaMemcpyAsync(dst, myStage, n * es, cudaMemcpyDeviceToDevice, stream));
  }
}


// ---- data movement ---------------------------------------------------------------------

namespace {

std::vector<size_t> prefix(const std::vector<size_t>& v) {
  std::vector<size_t> off(v.size() + 1, 0);
  for (size_t i = 0; i < v.size(); i++) off[i + 1] = off[i] + v[i];
  return off;
}

// A view of the pool's bulk region as a PeerBuffer-like target for staged calls.
struct StagedOut {
  PeerPtrs ptrs;
  void* mc;
  char* mine;
};

StagedOut stagedBulk(const PeerContext& pc, size_t needBytes, const char* what) {
  const auto l = layoutOf(pc);
  GLB_ENFORCE_LE(needBytes, l.bulkBytes, what, ": ", needBytes,
                 " bytes do not fit the staging pool; register the buffer (PeerContext::registerBuffer / "
                 "allocSymmetric) or raise GLB_CUDA_STAGE_MB");
  StagedOut s;
  s.ptrs = pc.stagePtrs(l.bulkOff);
  s.mc = pc.stageMc(l.bulkOff);
  s.mine = static_cast<char*>(s.ptrs.p[pc.rank]);
  return s;
}

// Store-only kernels (broadcast / gather / alltoall pushes) are light on registers, so two
// CTAs per SM stay co-resident: their cap is 2x the reduce kernels'.
int bwBlocks(const PeerContext& pc, size_t bytes) {
  size_t want = ceilDiv(std::max<size_t>(bytes / 16, 1), static_cast<size_t>(kThreads) * 4);
  int cap = std::min({tuning().copyBlocks, 2 * pc.maxBlocks(), kMaxBlocks});
  return std::max(1, static_cast<int>(std::min<size_t>(want, static_cast<size_t>(cap))));
}

}  // namespace

void broadcast(PeerContext& pc, const PeerBuffer& buf, size_t byteOffset, size_t bytes, int root,
               cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::broadcast");
  if (bytes == 0 || pc.size == 1) return;
  GLB_ENFORCE(root >= 0 && root < pc.size, "broadcast: invalid root ", root);
  GLB_ENFORCE_LE(byteOffset + bytes, buf.bytes, "broadcast range exceeds the registered buffer");
  DeviceGuard g(pc.device);
  const bool vec = buf.vectorOk && byteOffset % 16 == 0;
  int mode = 0;
  // multimem.st pays off when the root would otherwise send P-1 copies: P > 2.
  if (buf.mc != nullptr && vec && pc.size > 2 && bytes >= tuning().bcastDirectMaxBytes) {
    mode = 2;
  } else if (bytes > tuning().bcastDirectMaxBytes && pc.size > 2) {
    mode = 1;
  }
  {
    long forced = envInt("CUDA_BCAST_MODE", -1);
    if (forced >= 0 && forced <= 2 && (forced != 2 || (buf.mc != nullptr && vec))) mode = static_cast<int>(forced);
  }
  const int blocks = bwBlocks(pc, mode == 1 ? bytes / pc.size * 2 : bytes);
  pc.launchGuard();
  launchBroadcast(pc.comm(), buf.ptrsAt(byteOffset), buf.mc ? static_cast<char*>(buf.mc) + byteOffset : nullptr,
                  bytes, root, mode, vec, blocks, stream);
  checkLaunch("broadcast");
}

void broadcast(PeerContext& pc, void* ptr, size_t bytes, int root, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::broadcast(staged)");
  if (bytes == 0 || pc.size == 1) return;
  DeviceGuard g(pc.device);
  const auto l = layoutOf(pc);
  const size_t piece = l.bulkBytes / 16 * 16;
  PeerPtrs stage = pc.stagePtrs(l.bulkOff);
  char* mine = static_cast<char*>(stage.p[pc.rank]);
  for (size_t done = 0; done < bytes; done += piece) {
    const size_t n = std::min(piece, bytes - done);
    char* p = static_cast<char*>(ptr) + done;
    if (pc.rank == root) GLB_CUDA_CHECK(cudaMemcpyAsync(mine, p, n, cudaMemcpyDeviceToDevice, stream));
    int mode = 0;
    if (pc.nvlsAvailable() && pc.size > 2 && n >= tuning().bcastDirectMaxBytes) {
      mode = 2;
    } else if (n > tuning().bcastDirectMaxBytes && pc.size > 2) {
      mode = 1;
    }
    pc.launchGuard();
    launchBroadcast(pc.comm(), stage, pc.stageMc(l.bulkOff), n, root, mode, true,
                    bwBlocks(pc, mode == 1 ? n / pc.size * 2 : n), stream);
    checkLaunch("broadcast(staged)");
    if (pc.rank != root) GLB_CUDA_CHECK(cudaMemcpyAsync(p, mine, n, cudaMemcpyDeviceToDevice, stream));
  }
}

namespace {
void gatherCommon(PeerContext& pc, const void* in, const PeerPtrs& outs, void* mcOut, bool vecOut,
                  const std::vector<size_t>& bytesPerRank, int onlyDst, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::allgather/gather");
  GLB_ENFORCE_EQ(static_cast<int>(bytesPerRank.size()), pc.size, "need one byte count per rank");
  auto off = prefix(bytesPerRank);
  const bool vec = vecOut && reinterpret_cast<uintptr_t>(in) % 16 == 0;
  pc.launchGuard();
  launchGatherPush(pc.comm(), in, outs, mcOut, off.data(), bytesPerRank.data(), onlyDst, vec,
                   bwBlocks(pc, bytesPerRank[pc.rank]), stream);
  checkLaunch("allgather/gather");
}
}  // namespace

void allgatherv(PeerContext& pc, const void* in, const PeerBuffer& out, size_t outOffset,
                const std::vector<size_t>& bytesPerRank, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto off = prefix(bytesPerRank);
  GLB_ENFORCE_LE(outOffset + off.back(), out.bytes, "allgather output exceeds the registered buffer");
  if (pc.size == 1) {
    char* dst = static_cast<char*>(out.local) + outOffset;
    if (dst != in && off.back() > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(dst, in, off.back(), cudaMemcpyDeviceToDevice, stream));
    return;
  }
  // No multicast here: an allgather's bottleneck is what every GPU RECEIVES ((P-1)/P of the
  // output either way); multimem.st would only relieve the uplink and additionally deliver
  // each block back to its sender (measured at P=2: 312 GB/s with, 540 GB/s without).
  gatherCommon(pc, in, out.ptrsAt(outOffset), nullptr, out.vectorOk && outOffset % 16 == 0, bytesPerRank, -1, stream);
}

void allgatherv(PeerContext& pc, const void* in, void* out, const std::vector<size_t>& bytesPerRank,
                cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto off = prefix(bytesPerRank);
  if (pc.size == 1) {
    if (out != in && off.back() > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, off.back(), cudaMemcpyDeviceToDevice, stream));
    return;
  }
  auto st = stagedBulk(pc, off.back(), "allgather");
  gatherCommon(pc, in, st.ptrs, nullptr, true, bytesPerRank, -1, stream);
  if (off.back() > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, st.mine, off.back(), cudaMemcpyDeviceToDevice, stream));
}

void gatherv(PeerContext& pc, const void* in, const PeerBuffer& out, size_t outOffset,
             const std::vector<size_t>& bytesPerRank, int root, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  GLB_ENFORCE(root >= 0 && root < pc.size, "gather: invalid root ", root);
  if (pc.size == 1) {
    char* dst = static_cast<char*>(out.local) + outOffset;
    if (dst != in && bytesPerRank[0] > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(dst, in, bytesPerRank[0], cudaMemcpyDeviceToDevice, stream));
    return;
  }
  gatherCommon(pc, in, out.ptrsAt(outOffset), nullptr, out.vectorOk && outOffset % 16 == 0, bytesPerRank, root, stream);
}

void gatherv(PeerContext& pc, const void* in, void* out, const std::vector<size_t>& bytesPerRank, int root,
             cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto off = prefix(bytesPerRank);
  if (pc.size == 1) {
    if (out != in && off.back() > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, off.back(), cudaMemcpyDeviceToDevice, stream));
    return;
  }
  auto st = stagedBulk(pc, off.back(), "gather");
  gatherCommon(pc, in, st.ptrs, nullptr, true, bytesPerRank, root, stream);
  if (pc.rank == root && off.back() > 0) {
    GLB_CUDA_CHECK(cudaMemcpyAsync(out, st.mine, off.back(), cudaMemcpyDeviceToDevice, stream));
  }
}

namespace {
void alltoallCommon(PeerContext& pc, const void* in, const std::vector<size_t>& sendBytes, const PeerPtrs& outs,
                    bool vecOut, const std::vector<size_t>& recvBytes, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::alltoall");
  GLB_ENFORCE_EQ(static_cast<int>(sendBytes.size()), pc.size, "alltoall: need one send size per rank");
  GLB_ENFORCE_EQ(static_cast<int>(recvBytes.size()), pc.size, "alltoall: need one recv size per rank");
  auto soff = prefix(sendBytes);
  auto roff = prefix(recvBytes);
  // dstOff[j] = where my chunk lands in rank j's output = rank j's receive offset
  // for source `me`. For the uniform case that is me * chunk; for the v-variant the
  // kernel reads it from rank j's pad (published there before the first barrier).
  bool uniform = true;
  for (int i = 0; i < pc.size; i++) uniform = uniform && sendBytes[i] == sendBytes[0] && recvBytes[i] == sendBytes[0];
  std::vector<size_t> dstOff(pc.size);
  for (int j = 0; j < pc.size; j++) dstOff[j] = uniform ? static_cast<size_t>(pc.rank) * sendBytes[0] : ~size_t(0);
  const bool vec = vecOut && reinterpret_cast<uintptr_t>(in) % 16 == 0;
  size_t total = soff.back();
  pc.launchGuard();
  launchAlltoallPush(pc.comm(), in, outs, soff.data(), sendBytes.data(), dstOff.data(), uniform ? nullptr : roff.data(),
                     -1, vec, bwBlocks(pc, total), stream);
  checkLaunch("alltoall");
}
}  // namespace

void alltoallv(PeerContext& pc, const void* in, const std::vector<size_t>& sendBytes, const PeerBuffer& out,
               size_t outOffset, const std::vector<size_t>& recvBytes, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto roff = prefix(recvBytes);
  GLB_ENFORCE_LE(outOffset + roff.back(), out.bytes, "alltoall output exceeds the registered buffer");
  if (pc.size == 1) {
    char* dst = static_cast<char*>(out.local) + outOffset;
    if (dst != in && sendBytes[0] > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(dst, in, sendBytes[0], cudaMemcpyDeviceToDevice, stream));
    return;
  }
  alltoallCommon(pc, in, sendBytes, out.ptrsAt(outOffset), out.vectorOk && outOffset % 16 == 0, recvBytes, stream);
}

void alltoallv(PeerContext& pc, const void* in, const std::vector<size_t>& sendBytes, void* out,
               const std::vector<size_t>& recvBytes, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto roff = prefix(recvBytes);
  if (pc.size == 1) {
    if (out != in && sendBytes[0] > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, sendBytes[0], cudaMemcpyDeviceToDevice, stream));
    return;
  }
  // Every rank stages the same span so the (data-dependent) fit check cannot diverge.
  auto st = stagedBulk(pc, roff.back(), "alltoall");
  alltoallCommon(pc, in, sendBytes, st.ptrs, true, recvBytes, stream);
  if (roff.back() > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, st.mine, roff.back(), cudaMemcpyDeviceToDevice, stream));
}

void scatter(PeerContext& pc, const void* in, const PeerBuffer& out, size_t outOffset, size_t bytes, int root,
             cudaStream_t stream) {
  DeviceGuard g(pc.device);
  GLB_ENFORCE(root >= 0 && root < pc.size, "scatter: invalid root ", root);
  GLB_ENFORCE_LE(outOffset + bytes, out.bytes, "scatter output exceeds the registered buffer");
  if (pc.size == 1) {
    char* dst = static_cast<char*>(out.local) + outOffset;
    if (dst != in && bytes > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(dst, in, bytes, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  std::vector<size_t> soff(pc.size), slen(pc.size, bytes), doff(pc.size, 0);
  for (int j = 0; j < pc.size; j++) soff[j] = static_cast<size_t>(j) * bytes;
  const bool vec = out.vectorOk && outOffset % 16 == 0 && (pc.rank != root || reinterpret_cast<uintptr_t>(in) % 16 == 0);
  pc.launchGuard();
  launchAlltoallPush(pc.comm(), in, out.ptrsAt(outOffset), soff.data(), slen.data(), doff.data(), nullptr, root, vec,
                     bwBlocks(pc, bytes * pc.size), stream);
  checkLaunch("scatter");
}

void scatter(PeerContext& pc, const void* in, void* out, size_t bytes, int root, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  if (pc.size == 1) {
    if (out != in && bytes > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  auto st = stagedBulk(pc, bytes, "scatter");
  std::vector<size_t> soff(pc.size), slen(pc.size, bytes), doff(pc.size, 0);
  for (int j = 0; j < pc.size; j++) soff[j] = static_cast<size_t>(j) * bytes;
  pc.launchGuard();
  launchAlltoallPush(pc.comm(), in, st.ptrs, soff.data(), slen.data(), doff.data(), nullptr, root,
                     pc.rank != root || reinterpret_cast<uintptr_t>(in) % 16 == 0, bwBlocks(pc, bytes * pc.size), stream);
  checkLaunch("scatter(staged)");
  if (bytes > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, st.mine, bytes, cudaMemcpyDeviceToDevice, stream));
}

namespace {
void reducePullCommon(PeerContext& pc, const PeerPtrs& ins, void* mcIn, bool vecIn, void* out,
                      const std::vector<size_t>& counts, DataType dt, ReduceOp op, cudaStream_t stream) {
  GLB_TRACE_RANGE("glb::cuda::reduce_scatter/reduce");
  GLB_ENFORCE_EQ(static_cast<int>(counts.size()), pc.size, "need one element count per rank");
  GLB_ENFORCE(op != ReduceOp::CUSTOM, "custom reductions run on the host path only");
  auto off = prefix(counts);
  const size_t es = elementSize(dt);
  // In-switch reduction only for P > 2 (at P = 2 it doubles the uplink traffic).
  const bool useMc = mcIn != nullptr && pc.size > 2 && tuning().nvlsReduceScatter && nvlsSupports(dt, op) &&
                     counts[pc.rank] * es >= tuning().nvlsMinBytes;
  pc.launchGuard();
  launchReducePull(pc.comm(), ins, mcIn, out, off.data(), counts.data(), dt, op, vecIn, useMc,
                   blocksFor(pc, counts[pc.rank] * es / 16, 1, tuning().maxBlocks), stream);
  checkLaunch("reduce_scatter");
}
}  // namespace

void reduce_scatter(PeerContext& pc, const PeerBuffer& in, size_t inOffset, void* out,
                    const std::vector<size_t>& counts, DataType dt, ReduceOp op, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto off = prefix(counts);
  const size_t es = elementSize(dt);
  GLB_ENFORCE_LE(inOffset + off.back() * es, in.bytes, "reduce_scatter input exceeds the registered buffer");
  if (pc.size == 1) {
    const char* src = static_cast<const char*>(in.local) + inOffset;
    if (src != out && counts[0] > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, src, counts[0] * es, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  reducePullCommon(pc, in.ptrsAt(inOffset), in.mc ? static_cast<char*>(in.mc) + inOffset : nullptr,
                   in.vectorOk && inOffset % 16 == 0, out, counts, dt, op, stream);
}

void reduce_scatter(PeerContext& pc, const void* in, void* out, const std::vector<size_t>& counts, DataType dt,
                    ReduceOp op, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  auto off = prefix(counts);
  const size_t es = elementSize(dt);
  if (pc.size == 1) {
    if (in != out && counts[0] > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, counts[0] * es, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  auto st = stagedBulk(pc, off.back() * es, "reduce_scatter");
  if (off.back() > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(st.mine, in, off.back() * es, cudaMemcpyDeviceToDevice, stream));
  reducePullCommon(pc, st.ptrs, st.mc, true, out, counts, dt, op, stream);
}

void reduce(PeerContext& pc, const PeerBuffer& in, size_t inOffset, const PeerBuffer& out, size_t outOffset,
            size_t count, DataType dt, ReduceOp op, int root, cudaStream_t stream) {
  DeviceGuard g(pc.device);
  GLB_ENFORCE(root >= 0 && root < pc.size, "reduce: invalid root ", root);
  const size_t es = elementSize(dt);
  GLB_ENFORCE_LE(inOffset + count * es, in.bytes, "reduce input exceeds the registered buffer");
  if (pc.size == 1) {
    const char* src = static_cast<const char*>(in.local) + inOffset;
    char* dst = static_cast<char*>(out.local) + outOffset;
    if (src != dst && count > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(dst, src, count * es, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  GLB_ENFORCE(out.peer[root] != nullptr, "reduce: the root's output is not registered");
  // Each rank reduces an equal slice and stores it straight into the root's output.
  std::vector<size_t> counts(pc.size);
  for (int r = 0; r < pc.size; r++) counts[r] = count / pc.size + (static_cast<size_t>(r) < count % pc.size ? 1 : 0);
  auto off = prefix(counts);
  char* dst = static_cast<char*>(out.peer[root]) + outOffset + off[pc.rank] * es;
  reducePullCommon(pc, in.ptrsAt(inOffset), in.mc ? static_cast<char*>(in.mc) + inOffset : nullptr,
                   in.vectorOk && inOffset % 16 == 0, dst, counts, dt, op, stream);
}

void reduce(PeerContext& pc, const void* in, void* out, size_t count, DataType dt, ReduceOp op, int root,
            cudaStream_t stream) {
  DeviceGuard g(pc.device);
  const size_t es = elementSize(dt);
  if (pc.size == 1) {
    if (in != out && count > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(out, in, count * es, cudaMemcpyDeviceToDevice, stream));
    return;
  }
  // Stage inputs in the first half of the bulk region, collect results in the second.
  const auto l = layoutOf(pc);
  const size_t half = l.bulkBytes / 2 / 16 * 16;
  GLB_ENFORCE_LE(count * es, half, "reduce: payload does not fit the staging pool; register the buffers");
  PeerPtrs ins = pc.stagePtrs(l.bulkOff);
  PeerPtrs outs = pc.stagePtrs(l.bulkOff + half);
  char* mine = static_cast<char*>(ins.p[pc.rank]);
  if (count > 0) GLB_CUDA_CHECK(cudaMemcpyAsync(mine, in, count * es, cudaMemcpyDeviceToDevice, stream));
  std::vector<size_t> counts(pc.size);
  for (int r = 0; r < pc.size; r++) counts[r] = count / pc.size + (static_cast<size_t>(r) < count % pc.size ? 1 : 0);
  auto off = prefix(counts);
  char* dst = static_cast<char*>(outs.p[root]) + off[pc.rank] * es;
  reducePullCommon(pc, ins, pc.stageMc(l.bulkOff), true, dst, counts, dt, op, stream);
  if (pc.rank == root && count > 0) {
    GLB_CUDA_CHECK(cudaMemcpyAsync(out, static_cast<char*>(outs.p[root]), count * es, cudaMemcpyDeviceToDevice, stream));
  }
}

}  // namespace cuda
}  // namespace glb
